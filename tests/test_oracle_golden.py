"""CPU: the oracle restatements (oracle/ref_qe.py, oracle/ref_cg.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import golden_inputs as gi
from helpers import rel_l2

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


@pytest.fixture(scope='module')
def built(oracle_sht):
    return oracle_sht


def test_qe_fast_path_matches_reference(gold, built):
    from oracle import ref_qe
    q = gi.qe_case()
    for k in ['ptt', 'p_p', 'p']:
        G, C = ref_qe.qe(k, q['tlm1'], q['elm1'], q['blm1'], q['cls'], q['nside'], q['lmax_qlm'])
        assert rel_l2(G, gold['qe_dd_' + k]) < 1e-12
        assert rel_l2(C, gold['qe_dd_x' + k[1:]]) < 1e-12 or np.linalg.norm(gold['qe_dd_x' + k[1:]]) < 1e-10


def test_qe_two_leg_symmetrisation_matches_reference(gold, built):
    """ivfs1 != ivfs2: the reference averages the estimator with the legs swapped (qest.py:327-332)."""
    from oracle import ref_qe
    q = gi.qe_case()
    for k in ['ptt', 'p_p', 'p']:
        a = ref_qe.qe(k, q['tlm1'], q['elm1'], q['blm1'], q['cls'], q['nside'], q['lmax_qlm'],
                      tbar2=q['tlm2'], ebar2=q['elm2'], bbar2=q['blm2'])
        b = ref_qe.qe(k, q['tlm2'], q['elm2'], q['blm2'], q['cls'], q['nside'], q['lmax_qlm'],
                      tbar2=q['tlm1'], ebar2=q['elm1'], bbar2=q['blm1'])
        assert rel_l2(0.5 * (a[0] + b[0]), gold['qe_ds_' + k]) < 1e-12


def test_generic_qe_equals_fast_path_in_reference(gold):
    """The reference's own claim (qest.py:23): qe_eval gives the same estimator as the fast path."""
    for k in ['ptt', 'p_p']:
        assert rel_l2(gold['qe_gen_' + k], gold['qe_dd_' + k]) < 1e-10


def test_cg_operators_tt(gold, built):
    from oracle import ref_cg
    c = gi.cg_case()
    nf = ref_cg.ninv_tt(c['ninv_t'][0], c['transf'])
    assert rel_l2(ref_cg.fwd_tt(c['x_t'], c['cls']['tt'], nf), gold['tt_fwd']) < 1e-11
    assert rel_l2(nf.apply_alm(c['x_t']), gold['tt_apply_alm']) < 1e-11
    assert rel_l2(nf.calc_prep(c['tmap']), gold['tt_prep']) < 1e-11
    from oracle.healpy_shim.healpy import almxfl
    assert rel_l2(almxfl(c['x_t'], ref_cg.pre_diag_tt(c['cls']['tt'], nf)), gold['tt_prediag']) < 1e-12
    assert abs(ref_cg.dot_tt(c['x_t'], gold['tt_fwd']) - gold['tt_dot'][0]) < 1e-11 * abs(gold['tt_dot'][0])


def test_cg_operators_pp(gold, built):
    from oracle import ref_cg
    c = gi.cg_case()
    for tag, ninv in (('pp', c['ninv_p1']), ('pp3', c['ninv_p3'])):
        nf = ref_cg.ninv_pp([n[0] for n in ninv], c['transf'])
        fe, fb = ref_cg.fwd_pp(c['x_e'], c['x_b'], c['cls'], nf)
        assert rel_l2(fe, gold[tag + '_fwd_e']) < 1e-11 and rel_l2(fb, gold[tag + '_fwd_b']) < 1e-11
        pe, pb = nf.calc_prep(c['qmap'], c['umap'])
        assert rel_l2(pe, gold[tag + '_prep_e']) < 1e-11 and rel_l2(pb, gold[tag + '_prep_b']) < 1e-11
        d = ref_cg.dot_pp((c['x_e'], c['x_b']), (fe, fb))
        assert abs(d - gold[tag + '_dot'][0]) < 1e-11 * abs(gold[tag + '_dot'][0])


def test_pcg_iteration_count_and_trace(gold, built):
    """Same iteration count and residual trace as the reference's cd_solve with the diagonal preconditioner."""
    from oracle import ref_cg
    from oracle.healpy_shim.healpy import almxfl
    c = gi.cg_case()
    nf = ref_cg.ninv_tt(c['ninv_t'][0], c['transf'])
    b = nf.calc_prep(c['tmap'])
    filt = ref_cg.pre_diag_tt(c['cls']['tt'], nf)
    x, it, trace = ref_cg.pcg(b, lambda v: ref_cg.fwd_tt(v, c['cls']['tt'], nf), lambda r: almxfl(r, filt),
                              ref_cg.dot_tt, 1e-6)
    ref_trace = gold['tt_diag_trace']
    assert it == int(ref_trace[-1][1])
    eps = np.array([t[1] for t in trace])
    assert np.allclose(eps, ref_trace[:, 2], rtol=1e-6)
    sol = almxfl(x, ref_cg._cli(c['cls']['tt']))     # apply_fini (opfilt_tt.py:39-41)
    assert rel_l2(sol, gold['tt_diag_soltn']) < 1e-8


def test_template_marginalisation_matches_reference(built):
    """Pixel-space templates (opfilt_tt marge_maps, opfilt_pp marge_qmaps / marge_umaps): oracle vs the unmodified
    reference (tests/golden/make_golden_templates.py)."""
    from oracle import ref_cg
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_templates.npz'))
    c, t = gi.cg_case(), gi.template_case()
    nf = ref_cg.ninv_tt(c['ninv_t'][0], c['transf'], marge_maps=t['tmaps'])
    assert rel_l2(nf.Minv, g['tm_pinv']) < 1e-9
    assert rel_l2(nf.apply_map(c['tmap']), g['tm_apply_map']) < 1e-11
    assert rel_l2(ref_cg.fwd_tt(c['x_t'], c['cls']['tt'], nf), g['tm_fwd']) < 1e-11
    nf = ref_cg.ninv_tt(c['ninv_t'][0], c['transf'], marge_monopole=False, marge_dipole=False, marge_maps=t['tmaps'][:1])
    assert rel_l2(nf.apply_map(c['tmap']), g['tmonly_apply_map']) < 1e-11
    nfp = ref_cg.ninv_pp([c['ninv_p1'][0][0]], c['transf'], marge_qmaps=t['qmaps'], marge_umaps=t['umaps'])
    q, u = nfp.apply_map(c['qmap'], c['umap'])
    assert rel_l2(q, g['pm_apply_q']) < 1e-11 and rel_l2(u, g['pm_apply_u']) < 1e-11
    fe, fb = ref_cg.fwd_pp(c['x_e'], c['x_b'], c['cls'], nfp)
    assert rel_l2(fe, g['pm_fwd_e']) < 1e-11 and rel_l2(fb, g['pm_fwd_b']) < 1e-11


def test_joint_tp_operators_match_reference(built):
    """qcinv/opfilt_tp.py (joint T + P filter): oracle vs the unmodified reference (tests/golden/make_golden_tp.py)."""
    from oracle import ref_cg
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_tp.npz'))
    c, t = gi.cg_case(), gi.template_case()
    x = (c['x_t'], c['x_e'], c['x_b'])
    for tag, ninv, kw in (('tp2', [c['ninv_t'][0], c['ninv_p1'][0][0]], dict(marge_monopole=True, marge_dipole=True)),
                          ('tp4', [c['ninv_t'][0]] + [m[0] for m in c['ninv_p3']], dict(marge_maps_t=t['tmaps'][:1]))):
        nf = ref_cg.ninv_tp(ninv, c['transf'], **kw)
        f = ref_cg.fwd_tp(*x, c['cls'], nf)
        for a, k in zip(f, 'teb'):
            assert rel_l2(a, g['%s_fwd_%s' % (tag, k)]) < 1e-11
        for a, k in zip(nf.calc_prep(c['tmap'], c['qmap'], c['umap']), 'teb'):
            assert rel_l2(a, g['%s_prep_%s' % (tag, k)]) < 1e-11
        assert abs(ref_cg.dot_tp(x, f) - g[tag + '_dot'][0]) < 1e-11 * abs(g[tag + '_dot'][0])
        for a, k in zip(ref_cg.lmat3(ref_cg.pre_diag_tp(c['cls'], nf), *x), 'teb'):
            assert rel_l2(a, g['%s_prediag_%s' % (tag, k)]) < 1e-11


def test_kappa_filter_operators_match_reference(built):
    """opfilt_kk (the temperature operators with C_L^kk = (L(L+1)/2)^2 C_L^pp as the signal spectrum): oracle vs the
    unmodified reference (tests/golden/make_golden_kk.py)."""
    from oracle import ref_cg
    from oracle.healpy_shim.healpy import almxfl
    g = np.load(os.path.join(os.path.dirname(GOLD), 'reference_golden_kk.npz'))
    c = gi.cg_case()
    clpp = gi.kk_cls(c['lmax'])['pp']
    l = np.arange(c['lmax'] + 1, dtype=float)
    clkk = clpp * (0.5 * l * (l + 1)) ** 2           # opfilt_kk.py:25-33
    nf = ref_cg.ninv_tt(c['ninv_t'][0], c['transf'])
    fwd = ref_cg.fwd_tt(c['x_t'], clkk, nf)
    assert rel_l2(fwd, g['kk_fwd']) < 1e-11
    assert rel_l2(nf.calc_prep(c['tmap']), g['kk_prep']) < 1e-11
    assert rel_l2(almxfl(c['x_t'], ref_cg.pre_diag_tt(clkk, nf)), g['kk_prediag']) < 1e-12
    assert abs(ref_cg.dot_tt(c['x_t'], fwd) - g['kk_dot'][0]) < 1e-11 * abs(g['kk_dot'][0])
