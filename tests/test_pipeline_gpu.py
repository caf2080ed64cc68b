"""GPU: the reference's pipeline shape end to end -- parameter file -> sims -> filtering library -> QE library --
checked against the CPU oracle fed with the same simulated maps."""
import importlib.util
import os
import tempfile

import numpy as np
import pytest

from helpers import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_params(name, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'params', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_idealized_example_param_file(oracle_sht):
    from oracle import ref_qe
    from oracle.healpy_shim.healpy import almxfl
    with tempfile.TemporaryDirectory() as tmp:
        par = _load_params('idealized_example', {'PLENS': tmp, 'PLK_NSIDE': '64', 'PLK_LMAX_IVF': '96', 'PLK_LMAX_QLM': '128',
                                                 'PLK_NSIMS': '4'})
        G = par.qlms_dd.get_sim_qlm('p', 0)
        Gtt = par.qlms_dd.get_sim_qlm('ptt', 0)
        Gds = par.qlms_ds.get_sim_qlm('p_p', 1)
        # oracle on the same simulated maps
        def ivf(idx):
            tmap = par.sims.get_sim_tmap(idx)
            q, u = par.sims.get_sim_pmap(idx)
            t = almxfl(oracle_sht.map2alm(tmap, lmax=par.lmax_ivf), par.ftl * np.where(par.transf > 0, 1 / par.transf, 0))
            e, b = oracle_sht.map2alm_spin([q, u], 2, lmax=par.lmax_ivf)
            fac = np.where(par.transf > 0, 1 / par.transf, 0)
            return t, almxfl(e, par.fel * fac), almxfl(b, par.fbl * fac)
        cls = {k: par.cl_len[k] for k in ['tt', 'ee', 'bb', 'te']}
        t0, e0, b0 = ivf(0)
        assert rel_l2(par.ivfs.get_sim_tlm(0), t0) < 1e-10 and rel_l2(par.ivfs.get_sim_elm(0), e0) < 1e-10
        Gr, _ = ref_qe.qe('p', t0, e0, b0, cls, par.nside, par.lmax_qlm)
        assert rel_l2(G, Gr) < 1e-10
        Gr, _ = ref_qe.qe('ptt', t0, e0, b0, cls, par.nside, par.lmax_qlm)
        assert rel_l2(Gtt, Gr) < 1e-10
        # sim x data, symmetrised (data = index -1 -> simulation nsims)
        t1, e1, b1 = ivf(1)
        td, ed, bd = ivf(-1)
        a = ref_qe.qe('p_p', t1, e1, b1, cls, par.nside, par.lmax_qlm, tbar2=td, ebar2=ed, bbar2=bd)
        b = ref_qe.qe('p_p', td, ed, bd, cls, par.nside, par.lmax_qlm, tbar2=t1, ebar2=e1, bbar2=b1)
        assert rel_l2(Gds, 0.5 * (a[0] + b[0])) < 1e-10
        # downstream libraries of the parameter file: responses (GPU Wigner transforms), semi-analytical N0 from the
        # filtered maps' spectra, mean-field subtracted QE spectrum of the "data" -- for a Gaussian (unlensed) sky the
        # raw spectrum scatters around N0
        R = par.qresp_dd.get_response('ptt', 'p')
        n0 = par.nhl_dd.get_sim_nhl(-1, 'ptt', 'ptt')
        qcl = par.qcls_dd.get_sim_qcl('ptt', -1)
        assert R.shape == n0.shape == qcl.shape == (par.lmax_qlm + 1,)
        assert np.all(R[2:100] > 0) and np.all(n0[2:100] > 0) and np.all(qcl[2:] >= 0)
        band = slice(8, 64)
        assert 0.3 < np.sum(qcl[band]) / np.sum(n0[band]) < 3.0
        assert 0.3 < np.sum(n0[band] / R[band] ** 2) / np.sum(1.0 / R[band]) < 3.0     # filters are optimal: N0 ~ R
        # cached on disk under the reference's file names
        assert os.path.exists(os.path.join(tmp, 'temp', 'idealized_example', 'qlms_dd', 'sim_p_0000.fits'))
        assert os.path.exists(os.path.join(tmp, 'temp', 'idealized_example', 'ivfs', 'sim_0000_tlm.fits'))


def test_cg_filter_equals_isotropic_filter_on_full_sky():
    """Known answer 7 of SURVEY.md section 8c: on an unmasked homogeneous-noise sky the CG solution equals the
    isotropic filter  map2alm(d) / (C_l + N_l / b_l^2) / b_l  (filt_simple.py:397-400 vs filt_cinv.py:152-168)."""
    from plancklens_b200 import hp, utils
    from plancklens_b200.qcinv import cd_solve, multigrid, opfilt_tt, util_alm
    import golden_inputs as gi
    nside, lmax = 64, 128
    rng = np.random.default_rng(0)
    cls = gi.toy_cls(lmax)
    transf = hp.gauss_beam(np.deg2rad(1.0), lmax=lmax)
    npix = 12 * nside ** 2
    ninv = np.full(npix, 1.0 / 50.0)
    tmap = rng.standard_normal(npix) * 5
    nf = opfilt_tt.alm_filter_ninv(ninv, transf, marge_monopole=False, marge_dipole=False)
    descr = [[0, ["diag_cl"], lmax, nside, np.inf, 1.0e-9, cd_solve.tr_cg, cd_solve.cache_mem()]]
    chain = multigrid.multigrid_chain(opfilt_tt, descr, cls, nf)
    sol = util_alm.dalm.zeros(lmax)
    chain.solve(sol, tmap)
    nl = 4 * np.pi / npix / ninv[0]
    iso = hp.almxfl(hp.map2alm(tmap, lmax=lmax, iter=0), utils.cli(cls['tt'] + nl * utils.cli(transf ** 2)) * utils.cli(transf))
    ls = gi.alm_ls(lmax)
    sel = (ls >= 2) & (ls <= nside)        # band-limited part, where HEALPix quadrature is accurate
    assert rel_l2(sol.numpy()[sel], iso[sel]) < 2e-3


@pytest.mark.parametrize('key', ['ptt', 'p_p', 'p'])
def test_pipelined_eval_qlms_equals_eval_qlm(key, tmp_path):
    """The double-buffered generator (H2D / transforms / D2H overlapped over consecutive simulations) returns exactly
    what the one-at-a-time call returns, for pinned and for pageable host inputs."""
    import torch
    import bench
    import golden_inputs as gi
    from plancklens_b200 import qest
    q = gi.qe_case()
    rng = np.random.default_rng(3)
    sims = []
    for i in range(5):
        alms = [gi.rand_alm(rng, q['lmax'], 2) * sc for sc in (1e-2, 1e-1, 1e-1)]
        if i % 2 == 0:   # pinned
            alms = [torch.from_numpy(a).pin_memory().numpy() for a in alms]
        sims.append(alms)
    ivfs = bench.mem_ivfs(sims, q['cls'], q['nside'])
    lib = qest.library_sepTP(str(tmp_path / 'qlms'), ivfs, ivfs, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
    idxs = [0, 1, 2, 3, 4, 1]
    got = list(lib.eval_qlms(key, idxs))
    assert [g[0] for g in got] == idxs
    for idx, G, C in got:
        Gr, Cr = lib.eval_qlm(key, idx)
        assert np.array_equal(G, Gr) and np.array_equal(C, Cr)


def test_cinv_tp_joint_filter_pipeline(tmp_path):
    """filt_cinv.cinv_tp + library_cinv_jTP (joint T + P CG filter) end to end at the smallest size the reference's
    constructor accepts (nside 512, lmax 1024): the returned alms solve the normal equations to the requested eps,
    and feed a joint-filtered QE library."""
    import torch
    from plancklens_b200 import hp, qest, utils
    from plancklens_b200.filt import filt_cinv
    from plancklens_b200.qcinv import cd_solve, opfilt_tp, util_alm
    import golden_inputs as gi
    nside, lmax = 512, 1024
    npix = 12 * nside ** 2
    rng = np.random.default_rng(5)
    cls = utils.camb_clfile(os.path.join(ROOT, 'plancklens_b200', 'data', 'cls', 'FFP10_wdipole_lensedCls.dat'), lmax=lmax)
    cls = {k: cls[k] for k in ('tt', 'ee', 'bb', 'te')}
    transf = hp.gauss_beam(np.deg2rad(10. / 60.), lmax=lmax)
    z = gi.pix_z(nside)
    mask = (np.abs(z) > 0.2).astype(float)
    vamin2 = hp.nside2pixarea(nside, degrees=True) * 3600.
    ninv = [[np.array([vamin2 / 35. ** 2]), mask], [np.array([vamin2 / 55. ** 2]), mask]]
    # the reference's default chain (filt_cinv.py:398-405) one level shallower and with a smaller dense block
    descr = [[2, ["split(dense, 32, diag_cl)"], 256, 128, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
             [1, ["split(stage(2), 256, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
             [0, ["split(stage(1), 512, diag_cl)"], lmax, nside, np.inf, 1.0e-5, cd_solve.tr_cg, cd_solve.cache_mem()]]
    cinv = filt_cinv.cinv_tp(str(tmp_path / 'cinv_tp'), lmax, nside, cls, transf, ninv, marge_monopole=True,
                             marge_dipole=True, chain_descr=descr)
    # CMB-like data: Gaussian sky with the fiducial spectra through the beam + white noise at the filter's level
    sim = [hp.almxfl(gi.rand_alm(rng, lmax, 2) / np.sqrt(2.), np.sqrt(cls[k]) * transf) for k in ('tt', 'ee', 'bb')]
    vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60.
    tmap = hp.alm2map(sim[0], nside) + rng.standard_normal(npix) * 35. / vamin
    qmap, umap = hp.alm2map_spin([sim[1], sim[2]], nside, 2, lmax)
    qmap = qmap + rng.standard_normal(npix) * 55. / vamin
    umap = umap + rng.standard_normal(npix) * 55. / vamin
    tlm, elm, blm = cinv.apply_ivf([tmap, qmap, umap])
    assert cinv.chain.last_monitor.trace[-1][1] <= 1e-5 and cinv.chain.niter < 100, cinv.chain.niter
    # residual of the normal equations in the chain's own (rescaled) variables: A x = b with x = S (returned / rescal)
    nf, dl = cinv.chain.n_inv_filt, cinv.chain.s_cls
    b = opfilt_tp.calc_prep([tmap, qmap, umap], dl, nf)
    ivf = util_alm.teblm([util_alm.dalm.from_numpy(hp.almxfl(a, utils.cli(cinv.rescal_cl[k]))) for a, k in zip((tlm, elm, blm), 'teb')])
    smat = np.zeros((lmax + 1, 3, 3))
    for (i, j), k in {(0, 0): 'tt', (0, 1): 'te', (1, 1): 'ee', (2, 2): 'bb'}.items():
        smat[:, i, j] = smat[:, j, i] = dl[k][:lmax + 1]
    x = opfilt_tp._lmat3(smat).apply(ivf)                      # Wiener solution = S * inverse-variance filtered
    r = b - opfilt_tp.fwd_op(dl, nf)(x)
    dot = opfilt_tp.dot_op()
    assert np.sqrt(dot(r, r) / dot(b, b)) < 3e-5
    torch.cuda.synchronize()

    class _sims:
        def hashdict(self): return {'sims': 'tp_test'}
        def get_sim_tmap(self, idx): return tmap
        def get_sim_pmap(self, idx): return qmap, umap
    ivfs = filt_cinv.library_cinv_jTP(str(tmp_path / 'ivfs'), _sims(), cinv, cls)
    assert rel_l2(ivfs.get_sim_elm(0), elm) < 1e-12
    qlms = qest.library_jtTP(str(tmp_path / 'qlms'), ivfs, ivfs, nside, lmax_qlm=lmax)
    G = qlms.get_sim_qlm('p', 0)
    assert np.all(np.isfinite(G)) and np.any(G != 0)


@pytest.mark.parametrize("key", ['ptt', 'p_p', 'p'])
def test_qe_power_equals_semi_analytic_n0(key):
    """Known answer 8 of SURVEY.md section 8c (the idea of the reference's own test, tests/test_w.py:58-62, taken to the
    map level): for a Gaussian, unlensed sky the raw spectrum of the unnormalised estimator is its Gaussian noise bias,
    C_L^{qq} = N0_L, with N0 computed semi-analytically from the spectra of the same filtered maps by one-dimensional
    Wigner-d integrals (nhl.get_nhl).  The two sides share no code below the filtered alms -- five syntheses, pixel
    products and a spin-1 analysis on one, Gauss-Legendre quadrature of small-d matrices on the other -- so a wrong
    factor, sign or spin convention anywhere in the QE path shows up as a ratio far from one.  nside 512, lmax 1024:
    ~5e4 modes per band, realisation-dependent N0, so the ratio holds to a few per cent."""
    from plancklens_b200 import hp
    with tempfile.TemporaryDirectory() as tmp:
        par = _load_params('idealized_example', {'PLENS': tmp, 'PLK_NSIDE': '512', 'PLK_LMAX_IVF': '1024', 'PLK_LMAX_QLM': '1200',
                                                 'PLK_NSIMS': '2', 'PLK_DEVICE_SIMS': '1'})
        for k in (key, 'x' + key[1:]):              # gradient and curl estimators
            q = par.qlms_dd.get_sim_qlm(k, 0)
            cl = hp.alm2cl(q) / par.qlms_dd.fsky12
            n0 = par.nhl_dd.get_sim_nhl(0, k, k)
            assert n0.shape == cl.shape
            for lo, hi in ((10, 100), (100, 300), (300, 600), (600, 900)):
                w = 2 * np.arange(lo, hi) + 1.
                r = np.sum(w * cl[lo:hi]) / np.sum(w * n0[lo:hi])
                assert abs(r - 1.) < 0.05, (k, lo, hi, r)
