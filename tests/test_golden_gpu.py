"""GPU parity of the product's reference-API layer (qcinv operators, CG chains, qest) against golden vectors
produced by the unmodified reference (tests/golden/make_golden.py) and against the CPU oracle.

Tolerances: operators 1e-10 relative L2 (north_star), CG: identical top-level iteration count and the same eps
trace, solution within 1e-7 (the stop test is eps <= 1e-6)."""
import os
import tempfile

import numpy as np
import pytest

import golden_inputs as gi
from helpers import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


@pytest.fixture(scope='module')
def mods():
    from plancklens_b200.qcinv import cd_solve, multigrid, opfilt_pp, opfilt_tt, util_alm
    return dict(cd_solve=cd_solve, multigrid=multigrid, opfilt_pp=opfilt_pp, opfilt_tt=opfilt_tt, util_alm=util_alm)


def test_opfilt_tt_operators(gold, mods):
    ot, ua = mods['opfilt_tt'], mods['util_alm']
    c = gi.cg_case()
    nf = ot.alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
    x = ua.dalm.from_numpy(c['x_t'])
    fwd = ot.fwd_op(c['cls'], nf)
    y = fwd(x)
    assert rel_l2(y.numpy(), gold['tt_fwd']) < 1e-10
    assert rel_l2(x.numpy(), c['x_t']) == 0.0          # fwd_op must not modify its argument
    assert rel_l2(ot.calc_prep(c['tmap'], c['cls'], nf).numpy(), gold['tt_prep']) < 1e-10
    assert rel_l2(ot.pre_op_diag(c['cls'], nf)(x).numpy(), gold['tt_prediag']) < 1e-12
    d = ot.dot_op()(x, y)
    assert abs(d - gold['tt_dot'][0]) < 1e-10 * abs(gold['tt_dot'][0])
    a = c['x_t'].copy()
    nf.apply_alm(a)                                     # numpy in-place form of the reference
    assert rel_l2(a, gold['tt_apply_alm']) < 1e-10
    z = ua.dalm.zeros(c['lmax'])
    assert fwd(z) is z                                  # zero short-circuit (opfilt_tt.py:68)


def test_opfilt_pp_operators(gold, mods):
    op, ua = mods['opfilt_pp'], mods['util_alm']
    c = gi.cg_case()
    for tag, ninv in (('pp', c['ninv_p1']), ('pp3', c['ninv_p3'])):
        nf = op.alm_filter_ninv(ninv, c['transf'])
        x = ua.eblm([ua.dalm.from_numpy(c['x_e']), ua.dalm.from_numpy(c['x_b'])])
        fwd = op.fwd_op(c['cls'], nf)
        y = fwd(x)
        e, b = y.numpy()
        assert rel_l2(e, gold[tag + '_fwd_e']) < 1e-10 and rel_l2(b, gold[tag + '_fwd_b']) < 1e-10
        pe, pb = op.calc_prep([c['qmap'], c['umap']], c['cls'], nf).numpy()
        assert rel_l2(pe, gold[tag + '_prep_e']) < 1e-10 and rel_l2(pb, gold[tag + '_prep_b']) < 1e-10
        d = op.dot_op()(x, y)
        assert abs(d - gold[tag + '_dot'][0]) < 1e-10 * abs(gold[tag + '_dot'][0])
        if tag == 'pp':
            e, b = op.pre_op_diag(c['cls'], nf)(x).numpy()
            assert rel_l2(e, gold['pp_prediag_e']) < 1e-12 and rel_l2(b, gold['pp_prediag_b']) < 1e-12


def test_template_marginalisation_matches_reference(mods):
    """opfilt_tt marge_maps (+ monopole + dipole) and opfilt_pp marge_qmaps / marge_umaps on the GPU against the
    unmodified reference (tests/golden/reference_golden_templates.npz)."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_templates.npz'))
    c, t = gi.cg_case(), gi.template_case()
    ua = mods['util_alm']
    for tag, kw in (('tm', dict(marge_monopole=True, marge_dipole=True, marge_maps=t['tmaps'])),
                    ('tmonly', dict(marge_maps=t['tmaps'][:1]))):
        nf = mods['opfilt_tt'].alm_filter_ninv(c['ninv_t'], c['transf'], **kw)
        assert rel_l2(nf.Pt_Nn1_P_inv, g[tag + '_pinv']) < 1e-9
        m = c['tmap'].copy()
        nf.apply_map(m)
        assert rel_l2(m, g[tag + '_apply_map']) < 1e-11
        y = mods['opfilt_tt'].fwd_op(c['cls'], nf)(ua.dalm.from_numpy(c['x_t'])).numpy()
        assert rel_l2(y, g[tag + '_fwd']) < 1e-10
        assert rel_l2(mods['opfilt_tt'].calc_prep(c['tmap'], c['cls'], nf).numpy(), g[tag + '_prep']) < 1e-10
    nf = mods['opfilt_pp'].alm_filter_ninv(c['ninv_p1'], c['transf'], marge_qmaps=t['qmaps'], marge_umaps=t['umaps'])
    q, u = c['qmap'].copy(), c['umap'].copy()
    nf.apply_map([q, u])
    assert rel_l2(q, g['pm_apply_q']) < 1e-11 and rel_l2(u, g['pm_apply_u']) < 1e-11
    assert rel_l2(nf.tniti, g['pm_tniti']) < 1e-9
    x = ua.eblm([ua.dalm.from_numpy(c['x_e']), ua.dalm.from_numpy(c['x_b'])])
    e, b = mods['opfilt_pp'].fwd_op(c['cls'], nf)(x).numpy()
    assert rel_l2(e, g['pm_fwd_e']) < 1e-10 and rel_l2(b, g['pm_fwd_b']) < 1e-10


def _solve(mods, opfilt, descr, cls, nf, sol, data):
    chain = mods['multigrid'].multigrid_chain(opfilt, descr, cls, nf)
    chain.solve(sol, data)
    return chain


def test_cg_tt_same_iterations_as_reference(gold, mods):
    """Two-level multigrid chain with a dense coarse preconditioner: same iteration count, same eps trace."""
    c = gi.cg_case()
    nf = mods['opfilt_tt'].alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
    sol = np.zeros(gold['tt_soltn'].size, dtype=complex)
    chain = _solve(mods, mods['opfilt_tt'], gi.chain_descr_t(mods['cd_solve']), c['cls'], nf, sol, c['tmap'])
    ref = gold['tt_trace']
    assert chain.niter == int(ref[-1][1])
    eps = np.array([t[1] for t in chain.last_monitor.trace])
    assert np.allclose(eps, ref[:, 2], rtol=1e-5)
    assert rel_l2(sol, gold['tt_soltn']) < 1e-7


def test_cg_tt_diag_only(gold, mods):
    c = gi.cg_case()
    cd = mods['cd_solve']
    nf = mods['opfilt_tt'].alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
    descr = [[0, ["diag_cl"], c['lmax'], c['nside'], np.inf, 1.0e-6, cd.tr_cg, cd.cache_mem()]]
    sol = np.zeros(gold['tt_diag_soltn'].size, dtype=complex)
    chain = _solve(mods, mods['opfilt_tt'], descr, c['cls'], nf, sol, c['tmap'])
    ref = gold['tt_diag_trace']
    assert chain.niter == int(ref[-1][1])
    assert np.allclose(np.array([t[1] for t in chain.last_monitor.trace]), ref[:, 2], rtol=1e-5)
    assert rel_l2(sol, gold['tt_diag_soltn']) < 1e-7


def test_kappa_filter_matches_reference(mods):
    """opfilt_kk: operators at 1e-10 and the two-level chain with the reference's iteration count and eps trace
    (golden: tests/golden/make_golden_kk.py, unmodified reference)."""
    from plancklens_b200.qcinv import opfilt_kk
    g = np.load(os.path.join(os.path.dirname(GOLD), 'reference_golden_kk.npz'))
    c, ua = gi.cg_case(), mods['util_alm']
    s_cls = gi.kk_cls(c['lmax'])
    nf = opfilt_kk.alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
    assert np.allclose(nf.get_fkl(), nf.get_ftl()) and nf.nlev_fkl == nf.nlev_ftl
    x = ua.dalm.from_numpy(c['x_t'])
    fwd = opfilt_kk.fwd_op(s_cls, nf)
    y = fwd(x)
    assert rel_l2(y.numpy(), g['kk_fwd']) < 1e-10
    assert rel_l2(opfilt_kk.calc_prep(c['tmap'], s_cls, nf).numpy(), g['kk_prep']) < 1e-10
    assert rel_l2(opfilt_kk.pre_op_diag(s_cls, nf)(x).numpy(), g['kk_prediag']) < 1e-12
    d = opfilt_kk.dot_op()(x, y)
    assert abs(d - g['kk_dot'][0]) < 1e-10 * abs(g['kk_dot'][0])
    assert set(fwd.hashdict()) == {'clkk_inv', 'n_inv_filt'}
    sol = np.zeros(g['kk_soltn'].size, dtype=complex)
    chain = _solve(mods, opfilt_kk, gi.chain_descr_t(mods['cd_solve']), s_cls, nf, sol, c['tmap'])
    ref = g['kk_trace']
    assert chain.niter == int(ref[-1][1])
    assert np.allclose(np.array([t[1] for t in chain.last_monitor.trace]), ref[:, 2], rtol=1e-5)
    assert rel_l2(sol, g['kk_soltn']) < 1e-7


@pytest.mark.parametrize('pol', [False, True])
def test_cg_graph_and_device_scalar_stages_equal_host_loop(gold, mods, pol, monkeypatch):
    """The inner multigrid stages run three ways -- the reference's host loop (cd_solve, PLK_CG_FIXED=0), the
    device-scalar loop (cd_solve_fixed) launched eagerly (PLK_CG_GRAPH=0) and replayed as a CUDA graph -- and
    must give the same solution, iteration count and eps trace."""
    c = gi.cg_case()
    ua = mods['util_alm']
    res = []
    for fixed, graph in (('0', '0'), ('1', '0'), ('1', '1')):
        monkeypatch.setenv('PLK_CG_FIXED', fixed)
        monkeypatch.setenv('PLK_CG_DEVTOP', fixed)       # '0': the reference's host loop at the top level as well
        monkeypatch.setenv('PLK_CG_GRAPH', graph)
        if pol:
            nf = mods['opfilt_pp'].alm_filter_ninv(c['ninv_p1'], c['transf'])
            n = gold['pp_soltn_e'].size
            sol = ua.eblm([np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)])
            chain = _solve(mods, mods['opfilt_pp'], gi.chain_descr_p(mods['cd_solve']), c['cls'], nf, sol, [c['qmap'], c['umap']])
            v = np.concatenate([sol.elm, sol.blm])
        else:
            nf = mods['opfilt_tt'].alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
            sol = np.zeros(gold['tt_soltn'].size, dtype=complex)
            chain = _solve(mods, mods['opfilt_tt'], gi.chain_descr_t(mods['cd_solve']), c['cls'], nf, sol, c['tmap'])
            v = sol.copy()
        used_graph = any(type(op).__name__ == 'graphed_op' and op.graph is not None for op in chain.bstage.pre_ops)
        assert used_graph == (graph == '1')
        res.append((v, chain.niter, np.array([t[1] for t in chain.last_monitor.trace])))
    for v, niter, eps in res[1:]:
        assert niter == res[0][1]
        assert np.allclose(eps, res[0][2], rtol=1e-6)
        assert rel_l2(v, res[0][0]) < 1e-9


def test_cg_pp_same_iterations_as_reference(gold, mods):
    c = gi.cg_case()
    ua = mods['util_alm']
    nf = mods['opfilt_pp'].alm_filter_ninv(c['ninv_p1'], c['transf'])
    n = gold['pp_soltn_e'].size
    sol = ua.eblm([np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)])
    chain = _solve(mods, mods['opfilt_pp'], gi.chain_descr_p(mods['cd_solve']), c['cls'], nf, sol, [c['qmap'], c['umap']])
    ref = gold['pp_trace']
    assert chain.niter == int(ref[-1][1])
    assert np.allclose(np.array([t[1] for t in chain.last_monitor.trace]), ref[:, 2], rtol=1e-5)
    assert rel_l2(sol.elm, gold['pp_soltn_e']) < 1e-7 and rel_l2(sol.blm, gold['pp_soltn_b']) < 1e-7


def test_joint_tp_filter_matches_reference(mods):
    """qcinv/opfilt_tp.py on the GPU: operators, then a two-level multigrid solve with the dense TEB coarse
    preconditioner -- same iteration count, eps trace and solution as the unmodified reference."""
    from plancklens_b200.qcinv import opfilt_tp
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_tp.npz'))
    c, t = gi.cg_case(), gi.template_case()
    ua = mods['util_alm']
    x = lambda: ua.teblm([ua.dalm.from_numpy(c['x_t']), ua.dalm.from_numpy(c['x_e']), ua.dalm.from_numpy(c['x_b'])])
    for tag, ninv, kw in (('tp2', [c['ninv_t'][0], c['ninv_p1'][0][0]], dict(marge_monopole=True, marge_dipole=True)),
                          ('tp4', [c['ninv_t'][0]] + [m[0] for m in c['ninv_p3']], dict(marge_maps_t=t['tmaps'][:1]))):
        nf = opfilt_tp.alm_filter_ninv(ninv, c['transf'], **kw)
        fwd = opfilt_tp.fwd_op(c['cls'], nf)
        r = fwd(x())
        for a, k in zip(r.numpy(), 'teb'):
            assert rel_l2(a, g['%s_fwd_%s' % (tag, k)]) < 1e-10
        p = opfilt_tp.calc_prep([c['tmap'], c['qmap'], c['umap']], c['cls'], nf)
        for a, k in zip(p.numpy(), 'teb'):
            assert rel_l2(a, g['%s_prep_%s' % (tag, k)]) < 1e-10
        d = opfilt_tp.dot_op()(x(), r)
        assert abs(d - g[tag + '_dot'][0]) < 1e-10 * abs(g[tag + '_dot'][0])
        for a, k in zip(opfilt_tp.pre_op_diag(c['cls'], nf)(x()).numpy(), 'teb'):
            assert rel_l2(a, g['%s_prediag_%s' % (tag, k)]) < 1e-11
        if tag == 'tp2':
            n = g['tp2_soltn_t'].size
            sol = ua.teblm([np.zeros(n, dtype=complex), np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)])
            chain = _solve(mods, opfilt_tp, gi.chain_descr_tp(mods['cd_solve']), c['cls'], nf, sol, [c['tmap'], c['qmap'], c['umap']])
            ref = g['tp2_trace']
            assert chain.niter == int(ref[-1][1])
            assert np.allclose(np.array([tr[1] for tr in chain.last_monitor.trace]), ref[:, 2], rtol=1e-5)
            for a, k in zip((sol.tlm, sol.elm, sol.blm), 'teb'):
                assert rel_l2(a, g['tp2_soltn_' + k]) < 1e-7
            assert any(type(op).__name__ == 'graphed_op' and op.graph is not None for op in chain.bstage.pre_ops)


class _mem_ivfs:
    lib_dir = None

    def __init__(self, q, tag):
        self.q, self.tag = q, tag

    def hashdict(self):
        return {'tag': self.tag}

    def get_fmask(self):
        return np.ones(12 * self.q['nside'] ** 2)

    def get_sim_tlm(self, idx): return self.q['tlm' + self.tag].copy()
    def get_sim_elm(self, idx): return self.q['elm' + self.tag].copy()
    def get_sim_blm(self, idx): return self.q['blm' + self.tag].copy()

    def get_sim_tmliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_tlm(idx), self.q['cls']['tt'])

    def get_sim_emliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_elm(idx), self.q['cls']['ee'])

    def get_sim_bmliklm(self, idx):
        from plancklens_b200 import hp
        return hp.almxfl(self.get_sim_blm(idx), self.q['cls']['bb'])


@pytest.mark.parametrize("merge", [True, False])
def test_qest_library_matches_reference(gold, merge):
    """qest.library_sepTP.get_sim_qlm for 'ptt', 'p_p', 'p' (same-leg and two-leg symmetrised) vs the reference."""
    from plancklens_b200 import qest
    q = gi.qe_case()
    with tempfile.TemporaryDirectory() as tmp:
        iv1, iv2 = _mem_ivfs(q, '1'), _mem_ivfs(q, '2')
        dd = qest.library_sepTP(os.path.join(tmp, 'dd'), iv1, iv1, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
        ds = qest.library_sepTP(os.path.join(tmp, 'ds'), iv1, iv2, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
        dd.merge_analysis = ds.merge_analysis = merge
        for k in ['ptt', 'p_p', 'p']:
            assert rel_l2(dd.get_sim_qlm(k, 0), gold['qe_dd_' + k]) < 1e-10
            xk = 'x' + k[1:]
            ref = gold['qe_dd_' + xk]
            got = dd.get_sim_qlm(xk, 0)
            assert np.linalg.norm(got - ref) < 1e-10 * max(np.linalg.norm(ref), np.linalg.norm(gold['qe_dd_' + k]))
            assert rel_l2(ds.get_sim_qlm(k, 0), gold['qe_ds_' + k]) < 1e-10
        # qlm auto-spectra within 1e-8 (north_star)
        from plancklens_b200 import hp
        for k in ['ptt', 'p_p', 'p']:
            cl, ref = hp.alm2cl(dd.get_sim_qlm(k, 0)), hp.alm2cl(gold['qe_dd_' + k])
            assert np.max(np.abs(cl[2:] - ref[2:]) / ref[2:]) < 1e-8
        # mean field over two "sims" = plain average (reference: qest.py:239-243)
        mf = dd.get_sim_qlm_mf('ptt', np.array([0, 1]))
        assert rel_l2(mf, gold['qe_dd_ptt']) < 1e-10    # both indices map to the same in-memory sim


def test_qest_library_extra_keys_match_reference():
    """qest.library_sepTP.get_sim_qlm beyond the lensing fast path: point-source, noise-inhomogeneity, modulation and
    rotation estimators, single-pair lensing keys, their sums, bias hardening and the key remap -- identical legs and
    two different filtering libraries (symmetrised) -- vs the unmodified reference (make_golden_qest_keys.py)."""
    from plancklens_b200 import hp, qest
    g = np.load(os.path.join(os.path.dirname(GOLD), 'reference_golden_qest_keys.npz'))
    q = gi.qe_case()
    with tempfile.TemporaryDirectory() as tmp:
        iv1, iv2 = gi.idx_ivfs(q, hp), gi.idx_ivfs(q, hp, shift=3)
        resp = gi.toy_resplib(q['lmax_qlm'])
        dd = qest.library_sepTP(os.path.join(tmp, 'dd'), iv1, iv1, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'], resplib=resp)
        ds = qest.library_sepTP(os.path.join(tmp, 'ds'), iv1, iv2, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'], resplib=resp)
        scale = {tag: max(np.linalg.norm(g[tag + '_' + k]) for k in ('pte', 'pet', 'peb', 'pbe')) for tag in ('dd', 'ds')}
        for k in gi.QEST_EXTRA_KEYS:
            for tag, lib in (('dd', dd), ('ds', ds)):
                ref, got = g['%s_%s' % (tag, k)], lib.get_sim_qlm(k, 1)
                assert got.shape == ref.shape
                nrm = np.linalg.norm(ref)
                if nrm == 0.:      # 'ptb' with identical legs vanishes identically in the reference
                    assert np.linalg.norm(got) <= 1e-12 * scale[tag], (tag, k)
                else:
                    assert np.linalg.norm(got - ref) <= 1e-10 * nrm, (tag, k)
        assert rel_l2(dd.get_sim_qlm_mf('p_eb', [0, 1]), g['dd_mf_p_eb']) < 1e-10
        assert rel_l2(dd.get_sim_qlm_mf('ptt_bh_s', [0, 1]), g['dd_mf_ptt_bh_s']) < 1e-10
        assert dd.get_fundkeys(['p_tp', 'ptt_bh_s', 'p_eb', 'stt', 'x_te']) == list(g['fundkeys'])
        for name in ('sim_pte_0001.fits', 'sim_xte_0001.fits', 'sim_stt_0001.fits', 'sim_a_p_0001.fits', 'sim_f_0001.fits'):
            assert os.path.exists(os.path.join(tmp, 'dd', name)), name      # the reference's cache names
        with pytest.raises(AssertionError):
            dd.get_sim_qlm('dtt', 1)        # listed in keys_fund, never implemented (reference: qest.py:199)


def test_qlm_auto_spectra_match_reference(gold):
    """north_star tolerance: auto-spectra of the estimates within 1e-8 of the reference's, all three estimators,
    gradient and curl."""
    from plancklens_b200 import hp, qest
    import tempfile as _tf
    q = gi.qe_case()
    with _tf.TemporaryDirectory() as tmp:
        iv = _mem_ivfs(q, '1')
        lib = qest.library_sepTP(os.path.join(tmp, 'dd'), iv, iv, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
        for k in ['ptt', 'p_p', 'p']:
            G, C = lib.eval_qlm(k, 0)
            for mine, ref in ((G, gold['qe_dd_' + k]), (C, gold['qe_dd_x' + k[1:]])):
                cl, clr = hp.alm2cl(mine), hp.alm2cl(ref)
                sel = clr > 1e-30 * clr.max()
                assert np.max(np.abs(cl[sel] / clr[sel] - 1)) < 1e-8


def test_shts_seam_signatures(oracle_sht):
    """plancklens_b200.shts exposes the four functions of the reference seam with its signatures."""
    from plancklens_b200 import shts, utils_spin
    rng = np.random.default_rng(0)
    nside, lmax = 16, 32
    from helpers import rand_alm
    a = rand_alm(rng, lmax)
    m = shts.alm2map(a, nside)
    assert rel_l2(m, oracle_sht.alm2map(a, nside, lmax=lmax)) < 1e-11
    assert rel_l2(shts.map2alm(m, lmax, iter=0), oracle_sht.map2alm(m, lmax=lmax)) < 1e-11
    g, c = rand_alm(rng, lmax, 2), rand_alm(rng, lmax, 2)
    q, u = shts.alm2map_spin([g, c], nside, 2, lmax)
    rq, ru = oracle_sht.alm2map_spin([g, c], nside, 2, lmax)
    assert rel_l2(q, rq) < 1e-11 and rel_l2(u, ru) < 1e-11
    e, b = shts.map2alm_spin([q, u], 2, lmax)
    re, rb = oracle_sht.map2alm_spin([rq, ru], 2, lmax=lmax)
    assert rel_l2(e, re) < 1e-11 and rel_l2(b, rb) < 1e-11
    t0, zero = utils_spin.alm2map_spin([a, a * 0], nside, 0, lmax)       # spin 0: G^0 = -T
    assert rel_l2(t0, -m) < 1e-13 and zero == 0.
    with pytest.raises(TypeError):
        shts.alm2map(a[:-1], nside)


def test_generic_qe_eval_matches_reference(gold):
    """qest.eval_qe (qresp.get_qes -> utils_qe.qe_eval) for 'ptt' and 'p_p' vs the reference's own generic path, and
    the 'p' key vs the fast path (the equivalence the reference claims at qest.py:23)."""
    from plancklens_b200 import qest
    q = gi.qe_case()
    get_alm = lambda a: {'t': q['tlm1'], 'e': q['elm1'], 'b': q['blm1']}[a].copy()
    for k in ['ptt', 'p_p']:
        G, C = qest.eval_qe(k, q['lmax'], q['cls'], get_alm, q['nside'], q['lmax_qlm'], verbose=False)
        assert rel_l2(G, gold['qe_gen_' + k]) < 1e-10
        ref_c = gold['qe_gen_x' + k[1:]]
        assert np.linalg.norm(C - ref_c) < 1e-10 * np.linalg.norm(gold['qe_gen_' + k])
    G, C = qest.eval_qe('p', q['lmax'], q['cls'], get_alm, q['nside'], q['lmax_qlm'], verbose=False)
    assert rel_l2(G, gold['qe_dd_p']) < 1e-10
    # two different legs: symmetrised estimator equals the fast path's average of the two orderings
    get_alm2 = lambda a: {'t': q['tlm2'], 'e': q['elm2'], 'b': q['blm2']}[a].copy()
    G, C = qest.eval_qe('ptt', q['lmax'], q['cls'], get_alm, q['nside'], q['lmax_qlm'], verbose=False, get_alm2=get_alm2)
    assert rel_l2(G, gold['qe_ds_ptt']) < 1e-10


def test_qe_term_lists_match_reference_structure():
    """qresp.get_qes: number of terms and leg spins for the lensing keys (host logic only)."""
    from plancklens_b200 import qresp
    cls = gi.toy_cls(20)
    n = {k: len(qresp.get_qes(k, 20, cls)) for k in ['ptt', 'p_p', 'p', 'pee', 'p_eb']}
    assert n['ptt'] == 1 and n['p_p'] == 4 and n['p'] == 9, n
    for q in qresp.get_qes('p', 20, cls):
        assert q.leg_a.spin_ou + q.leg_b.spin_ou == 1


def test_qecl_spectra_match_reference():
    """qecl.library (mean-field subtracted QE spectra) against the unmodified reference:
    auto- and cross-spectra within 1e-8 (north_star tolerance for qlm auto-spectra)."""
    from plancklens_b200 import hp, qecl, qest, utils
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_qecl.npz'))
    q = gi.qe_case()
    with tempfile.TemporaryDirectory() as tmp:
        iv = gi.idx_ivfs(q, hp)
        lib = qest.library_sepTP(os.path.join(tmp, 'dd'), iv, iv, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
        qcl = qecl.library(os.path.join(tmp, 'qcl'), lib, lib, np.array([1, 2, 3, 4]))
        for k1, k2 in (('ptt', 'ptt'), ('p', 'p'), ('p_p', 'ptt'), ('x', 'x')):
            ref = g['qcl_%s_%s' % (k1, k2)]
            got = qcl.get_sim_qcl(k1, 0, k2=k2)
            assert np.max(np.abs(got - ref)) < 1e-8 * np.max(np.abs(ref)), (k1, k2)
        ref = g['qcl_p_dat']
        assert np.max(np.abs(qcl.get_sim_qcl('p', -1) - ref)) < 1e-8 * np.max(np.abs(ref))
        st = qcl.get_sim_stats_qcl('ptt', [0, 5, 6])
        assert st.N == 3 and st.mean().shape == ref.shape
        assert os.path.exists(os.path.join(tmp, 'qcl', 'sim_qcl_stats_ptt_ptt_%s.pk' % utils.mchash([0, 5, 6])))
        assert np.array_equal(qcl.get_sim_stats_qcl('ptt', [0, 5, 6]).mean(), st.mean())      # second call: from the cache
        # average of spectra libraries (reference: qecl.py:151-223)
        av = qecl.average(os.path.join(tmp, 'qclav'), [qcl, qcl])
        assert np.allclose(av.get_sim_qcl('ptt', 0), qcl.get_sim_qcl('ptt', 0), rtol=1e-15, atol=0)
        assert np.allclose(av.get_dat_qcl('p'), qcl.get_sim_qcl('p', -1), rtol=1e-15, atol=0)
        assert list(av.mc_sims_mf) == [1, 2, 3, 4] and av.get_lmaxqcl('ptt', 'ptt') == qcl.get_lmaxqcl('ptt', 'ptt')
        sa = av.get_sim_stats_qcl('ptt', [0, 5, 6])
        assert sa.N == 3 and np.allclose(sa.mean(), st.mean(), rtol=1e-14, atol=0) and np.all(sa.sigmas() >= 0)
