"""GPU parity of the config-3 filter libraries -- filt_cinv.cinv_t, cinv_p, library_cinv_sepTP with the reference's
DEFAULT multigrid chains (reference filt_cinv.py:113-116, :237-239) -- against goldens produced by the UNMODIFIED
reference (tests/golden/make_golden_cinv.py -> reference_golden_cinv.npz, reference_golden_cinv_deep.npz) at the smallest
size the constructors accept (nside 512, lmax 1024): masked sky, anisotropic noise, monopole + dipole marginalisation.

north_star: "CG converging in the same iteration count to the same eps"; solutions within 1e-7 (eps_min = 1e-5 bounds how
well either side knows the solution; both follow the same iterates, so they agree far better than that).
"""
import os

import numpy as np
import pytest

import golden_inputs as gi
from helpers import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CLPATH = os.path.join(os.path.dirname(HERE), 'plancklens_b200', 'data', 'cls', 'FFP10_wdipole_lensedCls.dat')


def _traced(chain, store):
    orig = chain.log

    def log(stage, it, eps, **kw):
        store.append((stage.depth, it, eps))
        return orig(stage, it, eps, **kw)
    chain.log = log


def _top(trace):
    return np.array([t for t in trace if t[0] == 0])


def _check_trace(got, ref, what):
    assert got.shape == ref.shape, (what, 'top-level iterations', got.shape[0] - 1, ref.shape[0] - 1)
    assert np.array_equal(got[:, 1], ref[:, 1]), what
    assert np.allclose(got[:, 2], ref[:, 2], rtol=1e-5, atol=0), (what, got[:, 2], ref[:, 2])


@pytest.fixture(scope="module")
def case(oracle_sht):
    return gi.cinv_case(oracle_sht.alm2map, oracle_sht.alm2map_spin, CLPATH)


@pytest.fixture(scope="module")
def libs(case, tmp_path_factory):
    from plancklens_b200.filt import filt_cinv
    tmp = str(tmp_path_factory.mktemp('cinv'))
    c = case
    cinv_t = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), c['lmax'], c['nside'], c['cls'], c['transf'], c['ninv_t'],
                              marge_monopole=True, marge_dipole=True, marge_maps=[])
    cinv_p = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), c['lmax'], c['nside'], c['cls'], c['transf'], c['ninv_p'])
    ivfs = filt_cinv.library_cinv_sepTP(os.path.join(tmp, 'ivfs'), gi.fixed_sim_lib(c), cinv_t, cinv_p, c['cls'])
    return cinv_t, cinv_p, ivfs


def test_inputs_are_the_goldens_inputs(case):
    """the maps the reference filtered and the maps filtered here are the same arrays (both from the CPU oracle)"""
    g = np.load(os.path.join(HERE, 'golden', 'reference_golden_cinv.npz'))
    assert case['mask'].sum() == g['mask_sum'][0]
    assert np.allclose([case['tmap'].sum(), np.abs(case['tmap']).sum()], g['tmap_sum'], rtol=1e-13, atol=0)
    assert np.allclose([case['qmap'].sum(), np.abs(case['qmap']).sum()], g['qmap_sum'], rtol=1e-13, atol=0)


def test_default_chains_are_the_references(libs):
    cinv_t, cinv_p, _ = libs
    dt = [(r[0], r[1], r[2], r[3], r[4], r[5]) for r in cinv_t.chain_descr]
    assert [r[0] for r in dt] == [3, 2, 1, 0] and [r[2:4] for r in dt] == [(256, 128), (512, 256), (1024, 512), (1024, 512)]
    assert dt[0][1][0].startswith('split(dense(') and dt[0][1][0].endswith('), 64, diag_cl)')
    assert [r[4] for r in dt] == [3, 3, 3, np.inf] and [r[5] for r in dt] == [0.0, 0.0, 0.0, 1.0e-5]
    dp = [(r[0], r[1], r[2], r[3], r[4], r[5]) for r in cinv_p.chain_descr]
    assert [r[0] for r in dp] == [2, 1, 0] and [r[2:4] for r in dp] == [(512, 256), (1024, 512), (1024, 512)]
    assert dp[0][1][0].endswith('), 32, diag_cl)') and dp[-1][5] == 1.0e-5


def test_library_cinv_sepTP_matches_reference(case, libs):
    g = np.load(os.path.join(HERE, 'golden', 'reference_golden_cinv.npz'))
    cinv_t, cinv_p, ivfs = libs
    lmax = case['lmax']
    # isotropic approximations and the mask the libraries derive from n_inv
    for name, got in (('ftl', ivfs.get_ftl()), ('fel', ivfs.get_fel()), ('fbl', ivfs.get_fbl()),
                      ('tal_t', ivfs.get_tal('t')), ('tal_e', ivfs.get_tal('e'))):
        assert np.allclose(got, g[name], rtol=1e-12, atol=0), name
    assert ivfs.get_fmask().sum() == g['fmask_sum'][0]

    tr_t, tr_p = [], []
    _traced(cinv_t.chain, tr_t)
    tlm = ivfs.get_sim_tlm(0)
    _traced(cinv_p.chain, tr_p)
    elm = ivfs.get_sim_elm(0)
    blm = ivfs.get_sim_blm(0)
    _check_trace(_top(tr_t), g['t_trace'], 'T')
    _check_trace(_top(tr_p), g['p_trace'], 'P')
    from plancklens_b200 import hp
    for name, alm in (('tlm', tlm), ('elm', elm), ('blm', blm)):
        assert rel_l2(gi.alm_sample(alm, lmax), g[name + '_sample']) < 1e-7, name
        assert abs(np.linalg.norm(alm) / g[name + '_norm'][0] - 1.) < 1e-7, name
        cl = hp.alm2cl(alm)
        big = g[name + '_cl'] > 1e-6 * g[name + '_cl'].max()
        assert np.allclose(cl[big], g[name + '_cl'][big], rtol=1e-6, atol=0), name
    assert rel_l2(gi.alm_sample(ivfs.get_sim_tmliklm(0), lmax), g['tmliklm_sample']) < 1e-7
    assert rel_l2(gi.alm_sample(ivfs.get_sim_emliklm(0), lmax), g['emliklm_sample']) < 1e-7
    # cached under the reference's file names
    for a in 'teb':
        assert os.path.exists(os.path.join(ivfs.lib_dir, 'sim_0000_%slm.fits' % a))

    # warm start through cinv_t.apply_ivf(tmap, soltn=...) (reference filt_cinv.py:196-203)
    tr2 = []
    _traced(cinv_t.chain, tr2)
    rc = cinv_t.rescal_cl
    start = hp.almxfl(hp.almxfl(tlm, np.where(rc > 0, 1. / np.where(rc > 0, rc, 1.), 0.)), case['cls']['tt'] * rc ** 2)
    tlm2 = cinv_t.apply_ivf(1.05 * case['tmap'], soltn=start)
    _check_trace(_top(tr2), g['t2_trace'], 'T warm start')
    assert rel_l2(gi.alm_sample(tlm2, lmax), g['tlm2_sample']) < 1e-7


def test_t_and_p_filters_side_by_side(case, libs, tmp_path, monkeypatch):
    """library_cinv_sepTP runs the T and the P solve of one simulation on two host threads and two streams once both
    chains are warm (filt_simple.library_sepTP._filter_tp_concurrent): same bits as one after the other, same iteration
    counts, and the next simulations keep agreeing (lanes: no shared plans or reduction scratch)."""
    import torch
    from plancklens_b200.filt import filt_cinv
    cinv_t, cinv_p, _ = libs
    sims = gi.fixed_sim_lib(case)
    seq = filt_cinv.library_cinv_sepTP(str(tmp_path / 'seq'), sims, cinv_t, cinv_p, case['cls'])
    con = filt_cinv.library_cinv_sepTP(str(tmp_path / 'con'), sims, cinv_t, cinv_p, case['cls'])
    monkeypatch.setenv('PLK_TP_CONCURRENT', '0')
    want = {}
    for idx in (3, 4, 5):
        want[idx] = [x.clone() for x in seq.get_sim_teblm_dev(idx)] + [cinv_t.chain.niter, cinv_p.chain.niter]
    assert not hasattr(seq, '_p_pool') and seq._tp_ready()
    monkeypatch.setenv('PLK_TP_CONCURRENT', '1')
    for idx in (3, 4, 5):
        t, e, b = con.get_sim_teblm_dev(idx)
        assert hasattr(con, '_p_pool')
        torch.cuda.synchronize()
        for got, ref, name in zip((t, e, b), want[idx], 'teb'):
            assert torch.equal(got, ref), (idx, name)
        assert [cinv_t.chain.niter, cinv_p.chain.niter] == want[idx][3:]
    con.flush()
    seq.flush()
    from plancklens_b200 import hp
    assert np.array_equal(hp.read_alm(os.path.join(con.lib_dir, 'sim_0004_elm.fits')), want[4][1].cpu().numpy())
    # working ahead: the lanes filter simulations 7 and 8 while the caller does something else on its own stream
    monkeypatch.setenv('PLK_TP_CONCURRENT', '1')
    con.prefetch_dev([7, 8, 4])                                  # 4 is in the device store: ignored
    assert sorted(con._pending) == [7, 8]
    busy = torch.ones(1 << 22, device='cuda')
    for _ in range(20):
        busy = busy * 1.0000001
    got7, got8 = con.get_sim_teblm_dev(7, 't'), con.get_sim_teblm_dev(8)     # asking for one field collects all three
    assert not con._pending and set(con._dev_store()[7]) == {'t', 'e', 'b'}
    assert con.cg_iterations[7] == {'T': cinv_t.chain.niter, 'P': cinv_p.chain.niter} or con.cg_iterations[7]['T'] > 0
    monkeypatch.setenv('PLK_TP_CONCURRENT', '0')
    seq.prefetch_dev([7, 8])                                     # switched off: a no-op
    assert not getattr(seq, '_pending', {})
    ref7, ref8 = seq.get_sim_teblm_dev(7), seq.get_sim_teblm_dev(8)
    assert torch.equal(got7[0], ref7[0]) and all(torch.equal(a, b) for a, b in zip(got8, ref8))
    assert con.cg_iterations[8] == seq.cg_iterations[8] and con.cg_iterations[7] == seq.cg_iterations[7]
    monkeypatch.setenv('PLK_TP_CONCURRENT', '1')
    # one field at a time still goes through the plain path
    t6, = con.get_sim_teblm_dev(6, 't')
    e6, b6 = con.get_sim_teblm_dev(6, 'eb')
    monkeypatch.setenv('PLK_TP_CONCURRENT', '0')
    ref6 = seq.get_sim_teblm_dev(6)
    assert torch.equal(t6, ref6[0]) and torch.equal(e6, ref6[1]) and torch.equal(b6, ref6[2])


def test_lanes_have_their_own_plans_and_scratch():
    import threading
    import torch
    from plancklens_b200 import _lib, sht
    from helpers import rand_alm
    p0 = sht.get_plan(32, 64)
    with sht.use_lane(1):
        assert sht.lane() == 1 and _lib.load().plk_get_lane() == 1
        p1 = sht.get_plan(32, 64)
        with sht.use_lane(2):
            assert sht.get_plan(32, 64) is not p1
        assert sht.lane() == 1 and _lib.load().plk_get_lane() == 1
    assert sht.lane() == 0 and _lib.load().plk_get_lane() == 0 and p1 is not p0 and sht.get_plan(32, 64) is p0
    assert _lib.load().plk_set_lane(sht.MAX_LANES) != 0
    # the lane is per thread, and dots issued from two lanes on two streams at once agree with the serial ones
    rng = np.random.default_rng(5)
    a = [sht.dev_alm(rand_alm(rng, 700)) for _ in range(4)]
    want = [sht.alm_dot_fused([a[i]], [a[i + 1]])[0:1].clone() for i in (0, 2)]
    torch.cuda.synchronize()
    out, seen = {}, {}

    def work(k):
        with sht.use_lane(k), torch.cuda.stream(torch.cuda.Stream()):
            seen[k] = (sht.lane(), _lib.load().plk_get_lane())
            for _ in range(300):
                r = sht.alm_dot_fused([a[2 * k]], [a[2 * k + 1]])
            torch.cuda.current_stream().synchronize()
            out[k] = r[0:1].clone()
    th = [threading.Thread(target=work, args=(k,)) for k in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert seen == {0: (0, 0), 1: (1, 1)}
    assert torch.equal(out[0], want[0]) and torch.equal(out[1], want[1])


def test_inner_stage_traces_match_reference(case, libs):
    """the reference logs every inner multigrid stage as well (1105 lines for one T solve); with the host-scalar path
    (PLK_CG_FIXED=0, PLK_CG_GRAPH=0: same kernels, step lengths read back like the reference's) the whole log agrees"""
    g = np.load(os.path.join(HERE, 'golden', 'reference_golden_cinv.npz'))
    from plancklens_b200.filt import filt_cinv
    c = case
    old = {k: os.environ.get(k) for k in ('PLK_CG_FIXED', 'PLK_CG_GRAPH')}
    os.environ.update({'PLK_CG_FIXED': '0', 'PLK_CG_GRAPH': '0'})
    try:
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            cinv_t = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), c['lmax'], c['nside'], c['cls'], c['transf'], c['ninv_t'],
                                      marge_monopole=True, marge_dipole=True, marge_maps=[])
            tr = []
            _traced(cinv_t.chain, tr)
            tlm = cinv_t.apply_ivf(c['tmap'])
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    tr = np.array(tr)
    ref = g['t_trace_all']
    assert tr.shape == ref.shape and np.array_equal(tr[:, :2], ref[:, :2])
    assert np.allclose(tr[:, 2], ref[:, 2], rtol=1e-4, atol=0)
    assert rel_l2(gi.alm_sample(tlm, c['lmax']), g['tlm_sample']) < 1e-7


def test_deep_case(oracle_sht, tmp_path):
    """3 / 4 uK-arcmin noise (signal dominated to lmax): 8 temperature and 22 polarization top-level iterations"""
    fn = os.path.join(HERE, 'golden', 'reference_golden_cinv_deep.npz')
    if not os.path.exists(fn):
        pytest.skip('deep golden not generated')
    g = np.load(fn)
    from plancklens_b200.filt import filt_cinv
    c = gi.cinv_case(oracle_sht.alm2map, oracle_sht.alm2map_spin, CLPATH, nlev_t=3., nlev_p=4.)
    assert np.allclose([c['tmap'].sum(), np.abs(c['tmap']).sum()], g['tmap_sum'], rtol=1e-13, atol=0)
    tmp = str(tmp_path)
    cinv_t = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), c['lmax'], c['nside'], c['cls'], c['transf'], c['ninv_t'],
                              marge_monopole=True, marge_dipole=True, marge_maps=[])
    cinv_p = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), c['lmax'], c['nside'], c['cls'], c['transf'], c['ninv_p'])
    tr_t, tr_p = [], []
    _traced(cinv_t.chain, tr_t)
    tlm = cinv_t.apply_ivf(c['tmap'])
    _traced(cinv_p.chain, tr_p)
    elm, blm = cinv_p.apply_ivf([c['qmap'], c['umap']])
    _check_trace(_top(tr_t), g['t_trace'], 'T deep')
    _check_trace(_top(tr_p), g['p_trace'], 'P deep')
    for name, alm in (('tlm', tlm), ('elm', elm), ('blm', blm)):
        assert rel_l2(gi.alm_sample(alm, c['lmax']), g[name + '_sample']) < 1e-7, name
    # the same polarization filter pushed to eps_min = 1e-7: more than 25 iterations, i.e. through the `roundoff = 25`
    # refresh of the residual (cd_solve.py:79-81) that no other case reaches
    fn = os.path.join(HERE, 'golden', 'reference_golden_cinv_refresh.npz')
    if not os.path.exists(fn):
        pytest.skip('refresh golden not generated')
    g = np.load(fn)
    from plancklens_b200.qcinv import cd_solve
    lmax, nside = c['lmax'], c['nside']
    descr = [[2, ["split(dense(), 32, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
             [1, ["split(stage(2),  512, diag_cl)"], 1024, 512, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
             [0, ["split(stage(1), 1024, diag_cl)"], lmax, nside, np.inf, 1.0e-7, cd_solve.tr_cg, cd_solve.cache_mem()]]
    cinv_p7 = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p7'), lmax, nside, c['cls'], c['transf'], c['ninv_p'], chain_descr=descr)
    tr = []
    _traced(cinv_p7.chain, tr)
    elm, blm = cinv_p7.apply_ivf([c['qmap'], c['umap']])
    assert g['p_trace'][-1][1] > 25
    _check_trace(_top(tr), g['p_trace'], 'P to 1e-7')
    assert rel_l2(gi.alm_sample(elm, lmax), g['elm_sample']) < 1e-7 and rel_l2(gi.alm_sample(blm, lmax), g['blm_sample']) < 1e-7


def test_anisofilt_example_param_file(oracle_sht):
    """params/anisofilt_example.py (the reference's masked-sky parameter file shape: sims -> cinv_t / cinv_p ->
    library_cinv_sepTP -> library_ftl -> qest.library_sepTP) at its smallest legal size.  Checked against the CPU oracle:
    the filtered alms solve the reference's normal equations to the requested eps (size-independent property of the
    CG, evaluated with the oracle's own operators), and the 'p' estimate equals the oracle QE of the same filtered alms."""
    import tempfile
    from oracle import ref_cg, ref_qe
    from oracle.healpy_shim.healpy import almxfl
    from test_pipeline_gpu import _load_params
    with tempfile.TemporaryDirectory() as tmp:
        par = _load_params('anisofilt_example', {'PLENS': tmp, 'PLK_NSIDE': '512', 'PLK_LMAX_IVF': '1024',
                                                 'PLK_LMAX_QLM': '1024', 'PLK_NSIMS': '2'})
        lmax = par.lmax_ivf
        G = par.qlms_dd.get_sim_qlm('p', 0)
        assert par.cinv_t.chain.niter >= 5 and par.cinv_p.chain.niter >= 2
        tlm, elm, blm = par.ivfs.get_sim_tlm(0), par.ivfs.get_sim_elm(0), par.ivfs.get_sim_blm(0)
        raw_t = par.ivfs_raw.get_sim_tlm(0)
        ls = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        assert np.all(tlm[ls < par.lmin_ivf] == 0) and np.array_equal(tlm[ls >= par.lmin_ivf], raw_t[ls >= par.lmin_ivf])
        # temperature normal equations with the oracle's operators: (C^-1 + B N^-1 B) x = B N^-1 d, x = C * tlm
        ninv = par.mask * par.vamin2 / par.nlev_t ** 2
        nf = ref_cg.ninv_tt(ninv, par.transf, marge_monopole=True, marge_dipole=True)
        tmap = np.asarray(par.sims.get_sim_tmap(0))
        d = nf.apply_map(tmap)
        b = almxfl(oracle_sht.map2alm(d, lmax=lmax, iter=0), par.transf * (d.size / (4 * np.pi)))
        x = almxfl(raw_t, par.cl_ivf['tt'])
        r = b - ref_cg.fwd_tt(x, par.cl_ivf['tt'], nf)
        # the chain works on rescaled unknowns (cinv_t rescal_cl ~ sqrt(l(l+1)/2pi), filt_cinv.py:84-92): its eps is the
        # norm of the residual in those units
        rc = par.cinv_t.rescal_cl
        ri = np.where(rc > 0, 1. / np.where(rc > 0, rc, 1.), 0.)
        eps = np.sqrt(ref_cg.dot_tt(almxfl(r, ri), almxfl(r, ri)) / ref_cg.dot_tt(almxfl(b, ri), almxfl(b, ri)))
        assert eps < 2e-5, eps
        cls = {k: par.cl_len[k] for k in ['tt', 'ee', 'bb', 'te']}
        Gr, _ = ref_qe.qe('p', tlm, elm, blm, cls, par.nside, par.lmax_qlm)
        assert rel_l2(G, Gr) < 1e-10
