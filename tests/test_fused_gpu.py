"""GPU: the fused CG kernels against the separate kernels / numpy they replace -- one-kernel dot products with the
step-length arithmetic (cd_solve.py:69-71, :95-99), the paired solution / residual update, the analysis with the
S^-1 x term folded in (opfilt_tt.py:67-73, opfilt_pp.py:51-55), the one-kernel monopole / dipole sums."""
import numpy as np
import pytest

from helpers import alm_dot, alm_ls, rand_alm, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lmax", [0, 1, 7, 64, 513, 2048])
def test_dot_fused(lmax):
    import torch
    from plancklens_b200 import sht
    rng = np.random.default_rng(lmax)
    a = [rand_alm(rng, lmax) for _ in range(3)]
    b = [rand_alm(rng, lmax) for _ in range(3)]
    da, db = [sht.dev_alm(x) for x in a], [sht.dev_alm(x) for x in b]
    for n, lmin in ((1, 0), (2, 2), (3, 0)):
        w = (alm_ls(lmax) >= lmin)
        want = sum(alm_dot(a[j] * w, b[j], lmax) for j in range(n))
        out = sht.alm_dot_fused(da[:n], db[:n], lmin=lmin).cpu().numpy()
        assert abs(out[0] - want) <= 1e-12 * max(1.0, sum(abs(alm_dot(np.abs(a[j]) * w, np.abs(b[j]), lmax)) for j in range(n)))
        num = torch.tensor([3.5], dtype=torch.float64, device='cuda')
        o2 = sht.alm_dot_fused(da[:n], db[:n], lmin=lmin, num=num, scale=2.0).cpu().numpy()
        assert o2[0] == out[0]                                      # fixed summation order: bit-reproducible
        if want != 0:
            assert np.isclose(o2[1], 2.0 * 3.5 / out[0], rtol=1e-15) and o2[2] == -o2[1]
        o3 = sht.alm_dot_fused(da[:n], db[:n], lmin=lmin, den=num, scale=-1.0).cpu().numpy()
        assert np.isclose(o3[1], -out[0] / 3.5, rtol=1e-15) and o3[2] == -o3[1]
    # zero divisors give zero, not NaN (an exactly-zero residual must not poison the solve)
    z = torch.zeros_like(da[0])
    one = torch.ones(1, dtype=torch.float64, device='cuda')
    o = sht.alm_dot_fused([z], [da[0]], num=one).cpu().numpy()
    assert o[0] == 0 and o[1] == 0 and o[2] == 0
    o = sht.alm_dot_fused([da[0]], [db[0]], den=torch.zeros(1, dtype=torch.float64, device='cuda')).cpu().numpy()
    assert o[1] == 0
    assert float(sht.scalar_ratio(one, torch.zeros(1, dtype=torch.float64, device='cuda')).item()) == 0.0
    # many launches back to back on one stream: the ticket counter is reset by each
    ref = sht.alm_dot_fused(da[:2], db[:2]).clone()
    for _ in range(20):
        assert torch.equal(sht.alm_dot_fused(da[:2], db[:2]), ref)


def test_axpy2():
    import torch
    from plancklens_b200 import sht
    rng = np.random.default_rng(3)
    lmax = 300
    x1, y1, x2, y2 = [rand_alm(rng, lmax) for _ in range(4)]
    d = [sht.dev_alm(v) for v in (x1, y1, x2, y2)]
    a = torch.tensor([0.37], dtype=torch.float64, device='cuda')
    sht.alm_axpy2(d[1], d[0], d[3], d[2], a)
    assert rel_l2(d[1].cpu().numpy(), y1 + 0.37 * x1) < 1e-15 and rel_l2(d[3].cpu().numpy(), y2 - 0.37 * x2) < 1e-15


@pytest.mark.parametrize("nside,lmax", [(32, 64), (128, 200)])
def test_analysis_with_additive_term(nside, lmax):
    from plancklens_b200 import sht
    rng = np.random.default_rng(nside)
    plan = sht.get_plan(nside, lmax)
    npix = 12 * nside ** 2
    fl = sht.dev_fl(rng.uniform(0.5, 1.5, lmax + 1), lmax)
    afl = sht.dev_fl(rng.uniform(0.5, 1.5, lmax + 1), lmax)
    afl2 = sht.dev_fl(rng.uniform(0.5, 1.5, lmax + 1), lmax)
    m1, m2 = sht.dev_map(rng.standard_normal(npix)), sht.dev_map(rng.standard_normal(npix))
    x1, x2 = sht.dev_alm(rand_alm(rng, lmax)), sht.dev_alm(rand_alm(rng, lmax))
    want = plan.map2alm(m1, fl=fl) + sht.almxfl(x1, afl)
    got = plan.map2alm_add(m1, fl, x1, afl)
    assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 1e-14
    for spin in (1, 2):
        g, c = plan.map2alm_spin(m1, m2, spin, flg=fl, flc=afl)
        wg, wc = g + sht.almxfl(x1, afl), c + sht.almxfl(x2, afl2)
        gg, gc = plan.map2alm_spin_add(m1, m2, spin, fl, afl, x1, afl, x2, afl2)
        assert rel_l2(gg.cpu().numpy(), wg.cpu().numpy()) < 1e-14 and rel_l2(gc.cpu().numpy(), wc.cpu().numpy()) < 1e-14


@pytest.mark.parametrize("nside", [1, 8, 64])
def test_modes_dot_one_kernel(nside):
    from oracle import ref_geom as rg
    from plancklens_b200 import sht
    rng = np.random.default_rng(nside)
    plan = sht.get_plan(nside, max(2 * nside, 2))
    npix = 12 * nside ** 2
    m, w = rng.standard_normal(npix), rng.uniform(0, 1, npix)
    theta, phi = rg.pix2ang(nside)
    modes = np.array([np.ones(npix), np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
    dm = sht.dev_map(m)
    for _ in range(3):     # repeated launches: ticket counter reset
        got = plan.modes_dot(dm.clone(), w=sht.dev_map(w)).cpu().numpy()
        assert np.max(np.abs(got - modes @ (m * w))) < 1e-12 * npix ** 0.5


@pytest.mark.parametrize("nside,lmax", [(8, 16), (64, 100), (256, 300)])
def test_analysis_of_pixel_programs(nside, lmax):
    """plk_map2alm_pix_dev: the ring kernel evaluates sum_k s_k a_k b_k per pixel instead of reading a product map
    (QE leg products, N^-1 multiply) -- same alm as analysing the materialised product"""
    import torch
    from plancklens_b200 import sht
    rng = np.random.default_rng(nside)
    plan = sht.get_plan(nside, lmax)
    npix = 12 * nside ** 2
    m = [sht.dev_map(rng.standard_normal(npix)) for _ in range(9)]
    fl = sht.dev_fl(rng.uniform(0.5, 1.5, lmax + 1), lmax)
    re = [(1.0, m[0], m[1]), (1.0, m[2], m[3]), (-1.0, m[0], m[4]), (-1.0, m[2], m[5]), (1.0, m[6], m[7])]
    im = [(1.0, m[0], m[3]), (-1.0, m[2], m[1]), (-1.0, m[2], m[4]), (1.0, m[0], m[5]), (0.5, m[6], None)]
    mre = m[0] * m[1] + m[2] * m[3] - m[0] * m[4] - m[2] * m[5] + m[6] * m[7]
    mim = m[0] * m[3] - m[2] * m[1] - m[2] * m[4] + m[0] * m[5] + 0.5 * m[6]
    for spin in (1, 2):
        g, c = plan.map2alm_spin_pix(re, im, spin, flg=fl, flc=fl)
        wg, wc = plan.map2alm_spin(mre, mim, spin, flg=fl, flc=fl)
        assert rel_l2(g.cpu().numpy(), wg.cpu().numpy()) < 1e-13 and rel_l2(c.cpu().numpy(), wc.cpu().numpy()) < 1e-13
    a = plan.map2alm_pix(re, fl=fl)
    assert rel_l2(a.cpu().numpy(), plan.map2alm(mre, fl=fl).cpu().numpy()) < 1e-13
    x1, x2 = sht.dev_alm(rand_alm(rng, lmax)), sht.dev_alm(rand_alm(rng, lmax))
    g, c = plan.map2alm_spin_pix(re[:1], im[:1], 2, flg=fl, flc=fl, addg=x1, aflg=fl, addc=x2, aflc=fl)
    wg, wc = plan.map2alm_spin_add(m[0] * m[1], m[0] * m[3], 2, fl, fl, x1, fl, x2, fl)
    assert rel_l2(g.cpu().numpy(), wg.cpu().numpy()) < 1e-13 and rel_l2(c.cpu().numpy(), wc.cpu().numpy()) < 1e-13


def test_qe_fused_products_equal_separate_kernels(monkeypatch):
    import golden_inputs as gi
    from plancklens_b200 import hp, qest, sht
    q = gi.qe_case()
    d = sht.dev_alm
    cls = q['cls']
    twf = hp.almxfl(q['tlm1'], cls['tt']) + hp.almxfl(q['elm1'], cls['te'])
    ewf = hp.almxfl(q['elm1'], cls['ee']) + hp.almxfl(q['tlm1'], cls['te'])
    bwf = hp.almxfl(q['blm1'], cls['bb'])
    args = [d(x) for x in (q['tlm1'], q['elm1'], q['blm1'], twf, ewf, bwf)]
    out = {}
    for fused in ('1', '0'):
        monkeypatch.setenv('PLK_QE_FUSED', fused)
        qe = qest.qe_device(q['nside'], q['lmax'], q['lmax_qlm'])
        out[fused] = [qe.p(*args), qe.ptt(args[0], args[3]), qe.p_p(args[1], args[2], args[4], args[5])]
    for a, b in zip(out['1'], out['0']):
        for x, y in zip(a, b):
            assert rel_l2(x.cpu().numpy(), y.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("lmax,lsplit", [(8, 8), (64, 20), (256, 64), (300, 0)])
def test_split_preconditioner_kernels(lmax, lsplit):
    """multigrid.pre_op_split with a diagonal high-l branch: `alm_splice_xfl` = almxfl then alm_splice, and the dense
    branch packs the l <= lsplit block of the longer vector exactly as it packs a truncated copy (multigrid.py:163-182)"""
    import torch
    from plancklens_b200 import _lib, hp, sht
    from plancklens_b200.qcinv import util_alm
    rng = np.random.default_rng(lmax + lsplit)
    lo, hi = rand_alm(rng, lsplit), rand_alm(rng, lmax)
    fl = rng.standard_normal(lmax + 1)
    got = sht.alm_splice_xfl(sht.dev_alm(lo), sht.dev_alm(hi), sht.dev_fl(fl, lmax), lsplit).cpu().numpy()
    want = util_alm.alm_splice(lo, hp.almxfl(hi, fl), lsplit)
    assert np.array_equal(got, want)
    r0 = torch.empty((lsplit + 1) ** 2, dtype=torch.float64, device='cuda')
    r1 = torch.empty_like(r0)
    cut = sht.dev_alm(util_alm.alm_copy(hi, lmax=lsplit))
    sht.check(_lib.load().plk_alm2rlm_dev(lsplit, sht._ptr(cut), sht._ptr(r0), sht._stream()))
    sht.check(_lib.load().plk_alm2rlm_from_dev(lsplit, lmax, sht._ptr(sht.dev_alm(hi)), sht._ptr(r1), sht._stream()))
    assert torch.equal(r0, r1)
    assert _lib.load().plk_alm2rlm_from_dev(lsplit + 1, lsplit, sht._ptr(cut), sht._ptr(r0), sht._stream()) != 0


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 200, 1089, 4225])
def test_dense_product(n):
    """plk_dense_matvec_dev (dense coarse preconditioner, dense.py:110-119): A x against numpy, bit-reproducible from call to call"""
    import torch
    from plancklens_b200.qcinv import dense
    rng = np.random.default_rng(n)
    a, x = rng.standard_normal((n, n)), rng.standard_normal(n)
    A, X = torch.from_numpy(a).cuda(), torch.from_numpy(x).cuda()
    y = dense._matvec(A, X)
    assert np.all(np.abs(y.cpu().numpy() - a @ x) <= 4e-16 * n ** 0.5 * (np.abs(a) @ np.abs(x)) + 1e-300)
    assert torch.equal(dense._matvec(A, X), y)
