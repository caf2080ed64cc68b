"""Host-side wrappers around the hot path (filter rescaling by m, harmonic-space filter and simulation libraries, sums
of simulation libraries, running statistics, real-harmonic packing) against the unmodified reference
(tests/golden/make_golden_wrappers.py).  Pure numpy except the pixel-space variant of cmb_maps_harmonicspace."""
import os

import numpy as np
import pytest

import golden_inputs as gi

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_wrappers.npz')


@pytest.fixture(scope='module')
def g():
    return np.load(GOLD)


def _eq(a, b, tol=1e-14):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)


def test_library_fml_and_alm_copy(g):
    from plancklens_b200 import hp
    from plancklens_b200.filt import filt_util
    q = gi.qe_case()
    w = gi.wrapper_case(q)
    iv = gi.idx_ivfs(q, hp)
    iv.get_ftl, iv.get_fel, iv.get_fbl = (lambda: w['ftl']), (lambda: w['fel']), (lambda: w['fbl'])
    fml = filt_util.library_fml(iv, w['lmax_cut'], w['fm_t'], w['fm_e'], w['fm_b'])
    for name in ('get_ftl', 'get_fel', 'get_fbl'):
        assert _eq(getattr(fml, name)(), g['fml_' + name])
    for name in ('tlm', 'elm', 'blm', 'tmliklm', 'emliklm', 'bmliklm'):
        assert _eq(getattr(fml, 'get_sim_' + name)(2), g['fml_' + name]), name
    assert set(fml.hashdict()) == {'ivfs', 'filt_t', 'filt_e', 'filt_b'}
    assert _eq(filt_util._alm_copy(q['tlm1'], None, q['lmax'] + 5, q['lmax'] + 2), g['alm_copy_up'], 0)
    assert _eq(filt_util._alm_copy(q['tlm1'], -1, q['lmax'] - 7, 9), g['alm_copy_dn'], 0)


def test_fullsky_filter_on_alm_simulations(g, tmp_path):
    from plancklens_b200.filt import filt_simple
    q = gi.qe_case()
    w = gi.wrapper_case(q)
    lib = filt_simple.library_fullsky_alms_sepTP(str(tmp_path / 'f'), gi.alm_sims(q),
                                                 {'t': w['transf'], 'e': w['transf'], 'b': w['transf'] ** 2},
                                                 q['cls'], w['ftl'], w['fel'], w['fbl'], cache=False)
    for name in ('tlm', 'elm', 'blm', 'tmliklm', 'emliklm'):
        assert _eq(getattr(lib, 'get_sim_' + name)(1), g['alms_' + name]), name
    assert _eq(lib.get_tal('b'), g['alms_tal_b']) and lib.get_fmask().shape == (1,)


def test_sums_of_simulation_libraries(g):
    from plancklens_b200.sims import utils as su
    q = gi.qe_case()
    add_sim = su.sim_lib_add_sim([gi.map_sims(q, 1.0), gi.map_sims(q, -0.3)], weights=[0.7, 2.0])
    add_dat = su.sim_lib_add_dat([gi.map_sims(q, 1.0), gi.map_sims(q, -0.3)])
    for tag, lib in (('add_sim', add_sim), ('add_dat', add_dat)):
        for idx in (-1, 2):
            assert _eq(lib.get_sim_tmap(idx), g['%s_t_%d' % (tag, idx)])
            qm, um = lib.get_sim_pmap(idx)
            assert _eq(qm, g['%s_q_%d' % (tag, idx)]) and _eq(um, g['%s_u_%d' % (tag, idx)])
    assert add_sim.hashdict()['lib'] == 'add_sim' and add_dat.hashdict()['lib'] == 'add_dat'


def _harmonic_sims(q, w, nside=None):
    from plancklens_b200.sims import maps
    tr = {'t': w['transf'], 'e': w['transf'], 'b': w['transf']}
    return maps.cmb_maps_harmonicspace(gi.alm_sims(q), tr, {'t': w['nl_t'], 'e': w['nl_p'], 'b': w['nl_p']},
                                       gi.fixed_phas(q), nside=nside)


def test_harmonic_space_simulations(g):
    q = gi.qe_case()
    hs = _harmonic_sims(q, gi.wrapper_case(q))
    assert _eq(hs.get_sim_tmap(1), g['hs_tlm'])
    e, b = hs.get_sim_pmap(1)
    assert _eq(e, g['hs_elm']) and _eq(b, g['hs_blm'])
    assert set(hs.hashdict()) == {'sims_cmb_len', 'phas', 'noiset', 'noisee', 'noiseb', 'transft', 'transfe', 'transfb'}


@pytest.mark.gpu
def test_harmonic_space_simulations_in_pixel_space(g):
    q = gi.qe_case()
    hs = _harmonic_sims(q, gi.wrapper_case(q), nside=8)
    assert _eq(hs.get_sim_tmap(1), g['hs_tmap'], 1e-11)
    qm, um = hs.get_sim_pmap(1)
    assert _eq(qm, g['hs_qmap'], 1e-11) and _eq(um, g['hs_umap'], 1e-11)


def test_running_statistics_and_real_harmonics(g):
    from plancklens_b200 import utils
    rows = g['stats_rows']
    st, st_nocov = utils.stats(6), utils.stats(6, docov=False)
    for r in rows:
        st.add(r)
        st_nocov.add(r)
    assert st.N == 20 and _eq(st.mean(), g['stats_mean']) and _eq(st.avg(), g['stats_mean'])
    assert _eq(st.cov(), g['stats_cov'], 1e-12) and _eq(st.sigmas(), g['stats_sig'], 1e-12)
    assert _eq(st.sigmas_on_mean(), g['stats_som'], 1e-12) and _eq(st.corrcoeffs(), g['stats_corr'], 1e-12)
    assert _eq(st.inverse(), g['stats_inv'], 1e-11)
    assert _eq([st.get_chisq(rows[3] * 0.5), st.get_chisq_pte(rows[3] * 0.5)], g['stats_chisq'], 1e-11)
    rb = st.rebin_that_nooverlap(np.arange(6.), np.array([0, 2, 4]), np.array([1, 3, 5]), weights=np.arange(1., 7.))
    assert _eq(rb.mean(), g['stats_rb_mean'], 1e-13) and _eq(rb.cov(), g['stats_rb_cov'], 1e-12)
    assert _eq(st_nocov.sigmas(), g['stats_sig'], 1e-11)          # extension: diagonal moments are always kept
    with pytest.raises(AssertionError):
        st_nocov.cov()
    q = gi.qe_case()
    rlm = utils.alm2rlm(q['tlm1'])
    assert _eq(rlm, g['rlm'], 0)
    back = utils.rlm2alm(rlm)
    ref = q['tlm1'].copy()
    ref[:q['lmax'] + 1] = ref[:q['lmax'] + 1].real
    assert _eq(back, ref, 1e-15)


def test_cachers(tmp_path):
    from plancklens_b200.helpers import cachers
    a = np.arange(5.)
    for c in (cachers.cacher_mem(), cachers.cacher_npy(str(tmp_path / 'n')), cachers.cacher_pk(str(tmp_path / 'p'))):
        assert not c.is_cached('x')
        c.cache('x', a)
        assert c.is_cached('x') and np.array_equal(c.load('x'), a)
    none = cachers.cacher_none()
    none.cache('x', a)
    assert not none.is_cached('x')
    with pytest.raises(AssertionError):
        none.load('x')
    with pytest.raises(AssertionError):
        cachers.cacher_npy(str(tmp_path / 'n')).load('missing')


def test_unlensed_cmb_extra_fields():
    from plancklens_b200.sims import cmbs, phas
    lmax = 20
    l = np.arange(lmax + 1, dtype=float)
    cls = {'tt': 1.0 / (1 + l) ** 2, 'ee': 0.1 / (1 + l) ** 2, 'te': 0.2 / (1 + l) ** 2, 'pp': 1e-3 / (1 + l) ** 4, 'pt': 1e-3 / (1 + l) ** 3}
    lib = cmbs.sims_cmb_unl(cls, phas.lib_phas(None, 3, lmax))
    assert lib.fields == ['p', 't', 'e']
    alms = lib.get_sim_alms(3)
    assert alms.shape == (3, (lmax + 1) * (lmax + 2) // 2)
    assert np.array_equal(alms[0], lib.get_sim_plm(3)) and np.array_equal(alms[1], lib.get_sim_tlm(3))
    with pytest.raises(AssertionError):
        lib.get_sim_olm(3)


def test_n1_library_surface_matches_reference(tmp_path):
    """n1.library_n1 around the (absent) flat-sky integrator: default nodes, key ordering, derived-estimator sums, L
    sampling and splining, both sqlite caches -- against the unmodified reference given the same stand-in integrator
    (tests/golden/make_golden_n1.py); and the behaviour without one."""
    from plancklens_b200.n1 import n1
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_n1.npz'))
    c = gi.n1_case()
    lib = n1.library_n1(str(tmp_path / 'n1'), c['cltt'], c['clte'], c['clee'], lmaxphi=c['lmaxphi'])
    assert np.array_equal(lib.lps, g['lps']) and not n1.HASN1F
    with pytest.raises(NotImplementedError):
        lib.get_n1('ptt', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'])
    calls = []

    def backend(*a):
        calls.append(a[0])
        return gi.fake_n1l(*a)
    lib.n1l = backend
    for kA, kB in gi.N1_PAIRS:
        got = lib.get_n1(kA, 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'], kB=kB, ftlB=c['ftlB'])
        assert _eq(got, g['n1_%s_%s' % (kA, kB)], 1e-13), (kA, kB)
    flat = lib.get_n1('ptt', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'] - 10, n1_flat=lambda ell: ell ** 2 * (ell + 1.) ** 2)
    assert _eq(flat, g['n1_flat'], 1e-12)
    # a second library on the same directory serves everything from the sqlite caches, without an integrator
    lib2 = n1.library_n1(str(tmp_path / 'n1'), c['cltt'], c['clte'], c['clee'], lmaxphi=c['lmaxphi'])
    got = lib2.get_n1('p', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'], kB='ptt', ftlB=c['ftlB'])
    assert _eq(got, g['n1_p_ptt'], 1e-13)
    n = len(calls)
    assert np.all(lib.get_n1('ptt', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'], ftlB=c['ftlB'], remove_only=True) == 0)
    assert len(calls) == n
    # remove_only drops the splined curve only; it is rebuilt from the per-multipole cache, still without an integrator
    again = lib2.get_n1('ptt', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'], ftlB=c['ftlB'])
    assert _eq(again, g['n1_ptt_ptt'], 1e-13)
    with pytest.raises(AssertionError):
        n1.library_n1(str(tmp_path / 'n1'), 2 * c['cltt'], c['clte'], c['clee'], lmaxphi=c['lmaxphi'])  # hash mismatch
