"""GPU parity: CUDA transforms (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerance: north_star asks for 1e-10 relative L2 in FP64 against the reference path; the oracle and the CUDA
kernels use different normalisations of the same recurrence and agree to ~1e-13, so the tests hold 1e-11.
"""
import numpy as np
import pytest

from helpers import alm_dot, alm_size, rand_alm, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-11

CASES = [(4, 11), (8, 23), (16, 40), (32, 95), (64, 128), (128, 300), (64, 64), (128, 100)]   # last two: lmax <= nside (no-alias ring path)


@pytest.fixture(scope="module")
def sht():
    from plancklens_b200 import sht as _s
    return _s


@pytest.mark.parametrize("nside,lmax", CASES)
@pytest.mark.parametrize("spin", [0, 1, 2, 3])
def test_synthesis_matches_oracle(sht, oracle_sht, nside, lmax, spin):
    rng = np.random.default_rng(100 * nside + spin)
    plan = sht.get_plan(nside, lmax)
    if spin == 0:
        a = rand_alm(rng, lmax)
        ref = oracle_sht.alm2map(a, nside, lmax=lmax)
        got = plan.alm2map_host(0, a)
        assert rel_l2(got, ref) < TOL
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        ref = oracle_sht.alm2map_spin([g, c], nside, spin, lmax)
        got = plan.alm2map_host(spin, g, c)
        assert rel_l2(got[0], ref[0]) < TOL and rel_l2(got[1], ref[1]) < TOL


@pytest.mark.parametrize("nside,lmax", CASES)
@pytest.mark.parametrize("spin", [0, 1, 2, 3])
def test_analysis_matches_oracle(sht, oracle_sht, nside, lmax, spin):
    rng = np.random.default_rng(200 * nside + spin)
    plan = sht.get_plan(nside, lmax)
    npix = 12 * nside ** 2
    if spin == 0:
        m = rng.standard_normal(npix)
        ref = oracle_sht.map2alm(m, lmax=lmax, iter=0)
        got = plan.map2alm_host(0, m)
        assert rel_l2(got, ref) < TOL
    else:
        m1, m2 = rng.standard_normal(npix), rng.standard_normal(npix)
        ref = oracle_sht.map2alm_spin([m1, m2], spin, lmax=lmax)
        got = plan.map2alm_host(spin, m1, m2)
        assert rel_l2(got[0], ref[0]) < TOL and rel_l2(got[1], ref[1]) < TOL


@pytest.mark.parametrize("spin", [0, 2])
def test_mid_size_with_ramp(sht, oracle_sht, spin):
    """nside 512 / lmax 1024: seeds below the double range near the poles (scaled ramp-up is exercised)."""
    nside, lmax = 512, 1024
    rng = np.random.default_rng(7 + spin)
    plan = sht.get_plan(nside, lmax)
    if spin == 0:
        a = rand_alm(rng, lmax)
        ref = oracle_sht.alm2map(a, nside, lmax=lmax)
        got = plan.alm2map_host(0, a)
        assert rel_l2(got, ref) < TOL
        back_ref = oracle_sht.map2alm(ref, lmax=lmax, iter=0)
        back = plan.map2alm_host(0, ref)
        assert rel_l2(back, back_ref) < TOL
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        ref = oracle_sht.alm2map_spin([g, c], nside, spin, lmax)
        got = plan.alm2map_host(spin, g, c)
        assert rel_l2(got[0], ref[0]) < TOL and rel_l2(got[1], ref[1]) < TOL
        back_ref = oracle_sht.map2alm_spin(ref, spin, lmax=lmax)
        back = plan.map2alm_host(spin, ref[0], ref[1])
        assert rel_l2(back[0], back_ref[0]) < TOL and rel_l2(back[1], back_ref[1]) < TOL


@pytest.mark.parametrize("spin", [0, 1, 2, 3])
def test_adjointness_full_size(sht, spin):
    """Size-independent property at BASELINE.json's full size (nside 2048, lmax 2048):
    <map2alm(m), a> = (4 pi / npix) <m, alm2map(a)>  (analysis is the exact adjoint of synthesis)."""
    import torch
    nside, lmax = 2048, 2048
    rng = np.random.default_rng(11 + spin)
    plan = sht.get_plan(nside, lmax)
    npix = 12 * nside ** 2
    w = 4 * np.pi / npix
    if spin == 0:
        a = rand_alm(rng, lmax)
        m = rng.standard_normal(npix)
        ya = plan.alm2map(sht.dev_alm(a)).cpu().numpy()
        am = plan.map2alm(sht.dev_map(m)).cpu().numpy()
        lhs = alm_dot(am, a, lmax)
        rhs = w * float(np.dot(m, ya))
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        m1, m2 = rng.standard_normal(npix), rng.standard_normal(npix)
        y = plan.alm2map_spin(sht.dev_alm(g), sht.dev_alm(c), spin)
        gm, cm = plan.map2alm_spin(sht.dev_map(m1), sht.dev_map(m2), spin)
        lhs = alm_dot(gm.cpu().numpy(), g, lmax) + alm_dot(cm.cpu().numpy(), c, lmax)
        rhs = w * float(np.dot(m1, y[0].cpu().numpy()) + np.dot(m2, y[1].cpu().numpy()))
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), abs(rhs), 1e-300)
    torch.cuda.synchronize()


def test_closed_forms(sht):
    """Known answers: a_00 = sqrt(4 pi) -> map == 1; dipole alm of template_removal.xyz_to_alm -> x.r."""
    nside, lmax = 64, 64
    plan = sht.get_plan(nside, lmax)
    a = np.zeros(alm_size(lmax), dtype=complex)
    a[0] = np.sqrt(4 * np.pi)
    m = plan.alm2map_host(0, a)
    assert np.max(np.abs(m - 1.0)) < 1e-13
    from oracle import ref_geom as rg
    theta, phi = rg.pix2ang(nside)
    xyz = np.array([0.3, -1.2, 0.7])
    a = np.zeros(alm_size(lmax), dtype=complex)
    a[1] = xyz[2] * np.sqrt(4 * np.pi / 3)
    a[lmax + 1] = (-xyz[0] + 1j * xyz[1]) * np.sqrt(2 * np.pi / 3)
    m = plan.alm2map_host(0, a)
    ref = xyz[0] * np.sin(theta) * np.cos(phi) + xyz[1] * np.sin(theta) * np.sin(phi) + xyz[2] * np.cos(theta)
    assert np.max(np.abs(m - ref)) < 1e-13


def test_ring_stage_nside4096_lmax5000(sht, oracle_sht):
    """BASELINE.json configs[4] size (nside 4096, mmax 5000): every ring length up to n = 16384, including the
    single-buffer Bluestein path (M = 8192) and rings that alias (2 mmax >= n), against the numpy ring FFTs."""
    import torch
    nside, lmax = 4096, 5000
    rng = np.random.default_rng(4096)
    plan = sht.get_plan(nside, lmax)
    X = np.zeros((plan.nring, plan.pitch), dtype=complex)
    X[:, :lmax + 1] = rng.standard_normal((plan.nring, lmax + 1)) + 1j * rng.standard_normal((plan.nring, lmax + 1))
    X[:, 0] = X[:, 0].real
    got = plan.ring_synth(torch.from_numpy(X).cuda()).cpu().numpy()
    ref = oracle_sht.phase2map(nside, X[:, :lmax + 1])
    assert rel_l2(got, ref) < 1e-12
    mp = rng.standard_normal(12 * nside ** 2)
    Xo = plan.ring_anal(sht.dev_map(mp)).cpu().numpy()
    ref = oracle_sht.map2phase(nside, mp, lmax) * (4 * np.pi / (12 * nside ** 2))
    assert rel_l2(Xo[:, :lmax + 1], ref) < 1e-12


@pytest.mark.parametrize("spin", [0, 2])
def test_adjointness_nside4096(sht, spin):
    """Size-independent property at configs[4] size (nside 4096, lmax 4000)."""
    import torch
    nside, lmax = 4096, 4000
    rng = np.random.default_rng(21 + spin)
    plan = sht.get_plan(nside, lmax)
    npix = 12 * nside ** 2
    w = 4 * np.pi / npix
    if spin == 0:
        a = rand_alm(rng, lmax)
        m = rng.standard_normal(npix)
        ya = plan.alm2map(sht.dev_alm(a)).cpu().numpy()
        am = plan.map2alm(sht.dev_map(m)).cpu().numpy()
        lhs = alm_dot(am, a, lmax)
        rhs = w * float(np.dot(m, ya))
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        m1, m2 = rng.standard_normal(npix), rng.standard_normal(npix)
        y = plan.alm2map_spin(sht.dev_alm(g), sht.dev_alm(c), spin)
        gm, cm = plan.map2alm_spin(sht.dev_map(m1), sht.dev_map(m2), spin)
        lhs = alm_dot(gm.cpu().numpy(), g, lmax) + alm_dot(cm.cpu().numpy(), c, lmax)
        rhs = w * float(np.dot(m1, y[0].cpu().numpy()) + np.dot(m2, y[1].cpu().numpy()))
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), abs(rhs), 1e-300)
    torch.cuda.synchronize()
    sht.clear_plans()


@pytest.mark.parametrize("nside,lmax", [(16, 40), (128, 300), (512, 1024)])
@pytest.mark.parametrize("spin", [1, 2, 3])
def test_gradient_only_synthesis(sht, nside, lmax, spin):
    """alm2map_spin with no curl input runs the gradient-only kernel (8 instead of 12 FMA per unit): same maps as
    the general kernel fed with an explicit zero curl."""
    import torch
    rng = np.random.default_rng(31 * nside + spin)
    plan = sht.get_plan(nside, lmax)
    g = sht.dev_alm(rand_alm(rng, lmax, spin))
    fl = sht.dev_fl(rng.standard_normal(lmax + 1), lmax)
    a = plan.alm2map_spin(g, None, spin, flg=fl)
    b = plan.alm2map_spin(g, torch.zeros_like(g), spin, flg=fl)
    for x, y in zip(a, b):
        assert float(torch.linalg.norm(x - y) / torch.linalg.norm(y)) < 1e-13


def test_map2alm_refinement_smoothing_and_mask_apodization():
    """healpy-shaped helpers above the transforms: map2alm(iter=k) (HEALPix map2alm_iter: k Jacobi refinement passes),
    hp.smoothing and the reference's utils.apodize_mask, against the oracle / the unmodified reference run on the
    oracle (tests/golden/make_golden_apo.py)."""
    import os
    import golden_inputs as gi
    from plancklens_b200 import hp, utils
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_apo.npz'))
    a = gi.apo_case()
    for it in (1, 3):
        got = hp.map2alm(a['map'], lmax=a['lmax'], iter=it)
        assert rel_l2(got, g['alm_iter%d' % it]) < 1e-11
    assert rel_l2(hp.map2alm(a['map'], lmax=a['lmax']), g['alm_iter3']) < 1e-11          # healpy's default is iter=3
    assert rel_l2(hp.map2alm(a['map'], lmax=a['lmax'], iter=0), g['alm_iter3']) > 1e-3   # and it matters
    for method in ('gaussian', 'hybrid'):
        got = utils.apodize_mask(a['mask'], sigma_arcmin=a['sigma_arcmin'], lmax=a['lmax'], method=method, cache_dir=None)
        assert rel_l2(got, g['apo_' + method]) < 1e-10, method
    with pytest.raises(ValueError):
        utils.apodize_mask(a['mask'], sigma_arcmin=a['sigma_arcmin'], lmax=a['lmax'], method='tophat', cache_dir=None)
    with pytest.raises(NotImplementedError):
        hp.map2alm(a['map'], lmax=a['lmax'], use_weights=True)
    # (tlm, elm, blm) triple, pol=True -> (T, Q, U), and back
    rng = np.random.default_rng(3)
    tlm, elm, blm = rand_alm(rng, a['lmax']), rand_alm(rng, a['lmax'], 2), rand_alm(rng, a['lmax'], 2)
    T, Q, U = hp.alm2map([tlm, elm, blm], a['nside'], pol=True)
    Q0, U0 = hp.alm2map_spin([elm, blm], a['nside'], 2, a['lmax'])
    assert np.array_equal(T, hp.alm2map(tlm, a['nside'])) and np.array_equal(Q, Q0) and np.array_equal(U, U0)
    # (T, Q, U) triple, pol=True: spin-0 and spin-2 analyses
    t, e, b = hp.map2alm([a['map'], a['mask'], a['map'] * a['mask']], lmax=a['lmax'], iter=0)
    e0, b0 = hp.map2alm_spin([a['mask'], a['map'] * a['mask']], 2, lmax=a['lmax'])
    assert np.array_equal(e, e0) and np.array_equal(b, b0) and np.array_equal(t, hp.map2alm(a['map'], lmax=a['lmax'], iter=0))
