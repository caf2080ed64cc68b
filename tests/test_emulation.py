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
