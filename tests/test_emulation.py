"""CPU: the CUDA kernels' arithmetic (same __host__ __device__ code, driven from host loops by tests/emul) against
the oracle: table generation, seeds with the scaled ramp-up, normalised recurrences, fold / Bluestein / FFT stages."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import alm_size, rand_alm, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'emul', 'libplk_emul.so')
vp = ctypes.c_void_p


@pytest.fixture(scope='module')
def emul():
    src = os.path.join(HERE, 'emul', 'emul.cu')
    csrc = os.path.join(HERE, '..', 'plancklens_b200', 'csrc')
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(('.h', '.cuh'))]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-O2', '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
                               '-Wno-deprecated-gpu-targets', '-o', SO, src])
    return ctypes.CDLL(SO)


@pytest.mark.parametrize("nbatch", [1, 2, 4])
@pytest.mark.parametrize("nside,lmax", [(4, 11), (8, 23), (16, 40), (32, 70), (64, 100), (16, 16), (32, 20), (64, 64), (64, 63)])
def test_ring_fft_stage(emul, oracle_sht, nside, lmax, nbatch):
    rng = np.random.default_rng(nside)
    nring, pitch = 4 * nside - 1, (lmax + 2) & ~1
    X = np.zeros((nring, pitch), dtype=complex)
    X[:, :lmax + 1] = rng.standard_normal((nring, lmax + 1)) + 1j * rng.standard_normal((nring, lmax + 1))
    X[:, 0] = X[:, 0].real
    out = np.zeros(12 * nside ** 2)
    emul.emul_ring_synth(nside, lmax, pitch, vp(X.ctypes.data), vp(out.ctypes.data), nbatch)
    assert rel_l2(out, oracle_sht.phase2map(nside, X[:, :lmax + 1])) < 1e-13
    mp = rng.standard_normal(12 * nside ** 2)
    Xo = np.zeros((nring, pitch), dtype=complex)
    emul.emul_ring_anal(nside, lmax, pitch, vp(mp.ctypes.data), vp(Xo.ctypes.data), nbatch)
    ref = oracle_sht.map2phase(nside, mp, lmax) * (4 * np.pi / (12 * nside ** 2))
    assert rel_l2(Xo[:, :lmax + 1], ref) < 1e-13


@pytest.mark.parametrize("nside,lmax", [(8, 23), (32, 70)])
@pytest.mark.parametrize("spin", [0, 1, 2, 3])
def test_legendre_stage(emul, oracle_sht, nside, lmax, spin):
    rng = np.random.default_rng(10 * nside + spin)
    nring, pitch = 4 * nside - 1, (lmax + 2) & ~1
    a1, a2 = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
    X1 = np.zeros((nring, pitch), dtype=complex)
    X2 = np.zeros((nring, pitch), dtype=complex)
    emul.emul_legendre_synth(nside, lmax, spin, vp(a1.ctypes.data), vp(a2.ctypes.data) if spin else None, pitch,
                             vp(X1.ctypes.data), vp(X2.ctypes.data))
    R1, R2 = oracle_sht.legendre_synth(nside, spin, lmax, lmax, a1, a2 if spin else None)
    assert rel_l2(X1[:, :lmax + 1], R1) < 1e-12
    if spin:
        assert rel_l2(X2[:, :lmax + 1], R2) < 1e-12
    Y1 = np.zeros((nring, pitch), dtype=complex)
    Y2 = np.zeros((nring, pitch), dtype=complex)
    Y1[:, :lmax + 1] = rng.standard_normal((nring, lmax + 1)) + 1j * rng.standard_normal((nring, lmax + 1))
    Y2[:, :lmax + 1] = rng.standard_normal((nring, lmax + 1)) + 1j * rng.standard_normal((nring, lmax + 1))
    o1 = np.zeros(alm_size(lmax), dtype=complex)
    o2 = np.zeros_like(o1)
    emul.emul_legendre_anal(nside, lmax, spin, pitch, vp(Y1.ctypes.data), vp(Y2.ctypes.data), vp(o1.ctypes.data), vp(o2.ctypes.data))
    G, C = oracle_sht.legendre_anal(nside, spin, lmax, lmax, Y1[:, :lmax + 1].copy(), Y2[:, :lmax + 1].copy() if spin else None)
    w = 4 * np.pi / (12 * nside ** 2)
    assert rel_l2(o1 * w, G) < 1e-12
    if spin:
        assert rel_l2(o2 * w, C) < 1e-12


def test_seeds_with_scaled_rampup(emul, oracle_sht):
    """nside 128 / lmax 383: sin^m(theta) underflows the double range on polar rings; seeds stay finite, even-aligned,
    above the start threshold, and the synthesis still matches the oracle."""
    nside, lmax, spin = 128, 383, 2
    npair = 2 * nside
    n = (lmax + 1) * npair
    ks = np.zeros(n, dtype=np.int32)
    s = [np.zeros(n) for _ in range(4)]
    emul.emul_seeds(nside, lmax, spin, vp(ks.ctypes.data), *[vp(x.ctypes.data) for x in s])
    ks = ks.reshape(lmax + 1, npair)
    K = lmax - np.maximum(np.arange(lmax + 1), spin) + 1
    active = ks < K[:, None]
    assert np.all(ks[active] % 2 == 0)
    assert np.mean(~active) > 0.1          # polar pairs are skipped at high m
    mag = np.maximum(np.abs(s[1]), np.abs(s[3])).reshape(lmax + 1, npair)
    assert np.all(np.isfinite(mag)) and np.all(mag[active] >= 2.0 ** -121)
    rng = np.random.default_rng(3)
    a1, a2 = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
    nring, pitch = 4 * nside - 1, (lmax + 2) & ~1
    X1 = np.zeros((nring, pitch), dtype=complex)
    X2 = np.zeros((nring, pitch), dtype=complex)
    emul.emul_legendre_synth(nside, lmax, spin, vp(a1.ctypes.data), vp(a2.ctypes.data), pitch, vp(X1.ctypes.data), vp(X2.ctypes.data))
    R1, R2 = oracle_sht.legendre_synth(nside, spin, lmax, lmax, a1, a2)
    assert rel_l2(X1[:, :lmax + 1], R1) < 1e-12 and rel_l2(X2[:, :lmax + 1], R2) < 1e-12


@pytest.mark.parametrize("nside_in,nside_out", [(1, 1), (2, 1), (8, 2), (16, 16), (32, 4), (64, 8), (128, 64), (64, 4), (128, 4), (256, 4)])
def test_udgrade_index_arithmetic(emul, nside_in, nside_out):
    """HEALPix ring <-> (face, x, y) arithmetic of the device degrade kernel against the oracle's NEST-based
    ud_grade (sum of children) -- same host-device code, driven from a host loop"""
    from oracle import ref_geom as rg
    emul.emul_hpx_roundtrip_errors.restype = ctypes.c_longlong
    assert emul.emul_hpx_roundtrip_errors(nside_in) == 0
    rng = np.random.default_rng(nside_in + nside_out)
    m = rng.standard_normal(12 * nside_in ** 2)
    out = np.zeros(12 * nside_out ** 2)
    emul.emul_udgrade_sum(nside_in, vp(m.ctypes.data), nside_out, vp(out.ctypes.data))
    ref = rg.ud_grade_sum(m, nside_out)
    # bit-identical: the kernel adds the children in numpy's own order (NEST order, pairwise summation)
    assert np.array_equal(out, ref)
    # which children belong to which parent is exact: an indicator map comes back as exact counts
    ones, unit = np.zeros(12 * nside_out ** 2), np.ones(12 * nside_in ** 2)
    emul.emul_udgrade_sum(nside_in, vp(unit.ctypes.data), nside_out, vp(ones.ctypes.data))
    assert np.all(ones == (nside_in // nside_out) ** 2)


def test_philox_host_device_code_matches_oracle(emul):
    from oracle import ref_rng
    for seed, stream, n in ((0, 0, 5), ((77 << 32) | 20000, (9 << 8) | 1, 1000)):
        out = np.zeros((n, 4), dtype=np.uint32)
        emul.emul_philox_words(ctypes.c_ulonglong(seed), ctypes.c_ulonglong(stream), ctypes.c_longlong(n), vp(out.ctypes.data))
        assert np.array_equal(out, ref_rng.words(seed, stream, n))
