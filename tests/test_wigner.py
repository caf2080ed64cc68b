"""Wigner small-d transforms: oracle against the Jacobi closed form and known answers
on the CPU; CUDA kernels against the oracle and the reference's own test identity (tests/test_w.py) on the GPU."""
import os

import numpy as np
import pytest

from helpers import rel_l2


def test_oracle_recurrence_matches_jacobi_closed_form():
    from oracle import ref_wigner as rw
    x = np.cos(np.linspace(0.05, np.pi - 0.05, 41))
    for s1, s2 in [(0, 0), (1, 0), (0, 1), (2, 0), (2, 2), (2, -2), (-2, 2), (1, -1), (3, 1), (-3, 2), (4, -2), (3, -3)]:
        d = rw.wigner_d_all(120, s1, s2, x)
        for l in (max(abs(s1), abs(s2)), 7, 30, 120):
            if l >= max(abs(s1), abs(s2)):
                ref = rw.wigner_d_jacobi(l, s1, s2, x)
                assert np.max(np.abs(d[l] - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref))), (s1, s2, l)


def test_oracle_known_answers():
    from oracle import ref_wigner as rw
    x, w = rw.get_xgwg(64)
    assert abs(w.sum() - 2.0) < 1e-13 and abs((w * x ** 2).sum() - 2.0 / 3.0) < 1e-13
    # d^l_00 = P_l ; d^1_11 = (1 + x)/2 ; d^1_10 = -sin(theta)/sqrt(2) ; d^2_22 = ((1+x)/2)^2
    d = rw.wigner_d_all(3, 0, 0, x)
    assert np.allclose(d[2], 0.5 * (3 * x ** 2 - 1), atol=1e-14)
    assert np.allclose(rw.wigner_d_all(1, 1, 1, x)[1], 0.5 * (1 + x), atol=1e-14)
    assert np.allclose(rw.wigner_d_all(1, 1, 0, x)[1], -np.sqrt(1 - x ** 2) / np.sqrt(2.), atol=1e-14)
    assert np.allclose(rw.wigner_d_all(2, 2, 2, x)[2], (0.5 * (1 + x)) ** 2, atol=1e-14)
    # orthogonality = round trip: wignercoeff(wignerpos(cl) w) = cl
    rng = np.random.default_rng(0)
    cl = rng.standard_normal(40)
    cl[:2] = 0
    back = rw.wignercoeff(rw.wignerpos(cl, x, 2, -2) * w, x, 2, -2, 39)
    assert rel_l2(back, cl) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("s1,s2", [(0, 0), (1, 0), (2, 2), (2, -2), (-1, 3), (0, -2), (3, 1), (4, -2)])
def test_gpu_wigner_transforms_match_oracle(s1, s2):
    from oracle import ref_wigner as rw
    from plancklens_b200 import wigners
    rng = np.random.default_rng(abs(s1) * 10 + abs(s2))
    lmax, n = 700, 1100
    xg, wg = wigners.get_xgwg(-1., 1., n)
    xr, wr = rw.get_xgwg(n)
    assert np.max(np.abs(xg - xr)) < 1e-14 and np.max(np.abs(wg - wr)) < 1e-15
    cl = rng.standard_normal(lmax + 1) / (1.0 + np.arange(lmax + 1)) ** 2
    xi = wigners.wignerpos(cl, xg, s1, s2)
    assert rel_l2(xi, rw.wignerpos(cl, xg, s1, s2)) < 1e-11
    f = rng.standard_normal(n) * wg
    assert rel_l2(wigners.wignercoeff(f, xg, s1, s2, lmax), rw.wignercoeff(f, xg, s1, s2, lmax)) < 1e-11


@pytest.mark.gpu
def test_gpu_wignerc_matches_oracle():
    from oracle import ref_wigner as rw
    from plancklens_b200 import utils_spin as us
    rng = np.random.default_rng(3)
    cl1 = rng.standard_normal(301) / (1.0 + np.arange(301)) ** 1.5
    cl2 = (rng.standard_normal(257) + 1j * rng.standard_normal(257)) / (1.0 + np.arange(257)) ** 1.5
    for args in [(0, 0, 0, 0), (1, 0, -1, 2), (2, -2, 1, 1), (-3, 2, 1, 0)]:
        got = us.wignerc(cl1, cl2, *args, lmax_out=400)
        assert rel_l2(got, rw.wignerc(cl1, cl2, *args, lmax_out=400)) < 1e-10


@pytest.mark.gpu
def test_w():
    """Port of the reference's own (and only) test, tests/test_w.py: for optimally filtered maps the analytical N0
    (nhl.get_nhl) equals the response (qresp.get_response), for the lensing ('p') and modulation ('f') estimators,
    separately and jointly filtered.  Same inputs, same tolerances (rtol 1e-6)."""
    from plancklens_b200 import hp, nhl, qresp, utils
    cls_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'plancklens_b200', 'data', 'cls')
    lmax_ivf, lmin_ivf = 500, 100
    nlev_t, nlev_p, beam_fwhm = 35., 35. * np.sqrt(2.), 6.
    lmax_qlm = lmax_ivf
    for ksource in ['p', 'f']:
        qe_keys = [ksource + 'tt', ksource + '_p', ksource]
        transf = hp.gauss_beam(beam_fwhm / 60. / 180. * np.pi, lmax=lmax_ivf)
        cls_len = utils.camb_clfile(os.path.join(cls_path, 'FFP10_wdipole_lensedCls.dat'))
        cls_weight = utils.camb_clfile(os.path.join(cls_path, 'FFP10_wdipole_lensedCls.dat'))
        nl = lambda nlev: (nlev / 60. / 180. * np.pi) ** 2 / transf ** 2
        fal_sepTP = {'tt': utils.cli(cls_len['tt'][:lmax_ivf + 1] + nl(nlev_t)),
                     'ee': utils.cli(cls_len['ee'][:lmax_ivf + 1] + nl(nlev_p)),
                     'bb': utils.cli(cls_len['bb'][:lmax_ivf + 1] + nl(nlev_p))}
        cls_ivfs_sepTP = {'tt': fal_sepTP['tt'].copy(), 'ee': fal_sepTP['ee'].copy(), 'bb': fal_sepTP['bb'].copy(),
                          'te': cls_len['te'][:lmax_ivf + 1] * fal_sepTP['tt'] * fal_sepTP['ee']}
        cls_dat = {'tt': cls_len['tt'][:lmax_ivf + 1] + nl(nlev_t), 'ee': cls_len['ee'][:lmax_ivf + 1] + nl(nlev_p),
                   'bb': cls_len['bb'][:lmax_ivf + 1] + nl(nlev_p), 'te': np.copy(cls_len['te'][:lmax_ivf + 1])}
        fal_jtTP = utils.cl_inverse(cls_dat)
        cls_ivfs_jtTP = utils.cl_inverse(cls_dat)
        for cls in [fal_sepTP, fal_jtTP, cls_ivfs_sepTP, cls_ivfs_jtTP]:
            for cl in cls.values():
                cl[:max(1, lmin_ivf)] *= 0.
        for qe_key in qe_keys:
            NG, NC, NGC, NCG = nhl.get_nhl(qe_key, qe_key, cls_weight, cls_ivfs_sepTP, lmax_ivf, lmax_ivf, lmax_out=lmax_qlm)
            RG, RC, RGC, RCG = qresp.get_response(qe_key, lmax_ivf, ksource, cls_weight, cls_len, fal_sepTP, lmax_qlm=lmax_qlm)
            if qe_key[1:] in ['tt', '_p']:
                assert np.allclose(NG[1:], RG[1:], rtol=1e-6), qe_key
                assert np.allclose(NC[2:], RC[2:], rtol=1e-6), qe_key
            assert np.all(NCG == 0.) and np.all(NGC == 0.)
            assert np.all(RCG == 0.) and np.all(RGC == 0.)
        NG, NC, NGC, NCG = nhl.get_nhl(ksource, ksource, cls_weight, cls_ivfs_jtTP, lmax_ivf, lmax_ivf, lmax_out=lmax_qlm)
        RG, RC, RGC, RCG = qresp.get_response(ksource, lmax_ivf, ksource, cls_weight, cls_len, fal_jtTP, lmax_qlm=lmax_qlm)
        assert np.allclose(NG[1:], RG[1:], rtol=1e-6), ksource
        assert np.allclose(NC[2:], RC[2:], rtol=1e-6), ksource
        assert np.all(NCG == 0.) and np.all(NGC == 0.)
        assert np.all(RCG == 0.) and np.all(RGC == 0.)


# ------------------------------------------------------------------------------------------------ responses and N0
RESP_KEYS = [('ptt', 'p', 'fal_sep'), ('p_p', 'p', 'fal_sep'), ('p', 'p', 'fal_jt'), ('x', 'x', 'fal_jt'),
             ('ftt', 'f', 'fal_sep'), ('ptt', 'f', 'fal_sep'), ('p', 'p', 'fal_tb'), ('ptt_bh_f', 'p', 'fal_sep')]
NHL_KEYS = [('ptt', 'ptt', 'cls_ivfs_sep'), ('p_p', 'p_p', 'cls_ivfs_sep'), ('p', 'p', 'cls_ivfs_jt'),
            ('ptt', 'p_p', 'cls_ivfs_sep'), ('p', 'p', 'cls_ivfs_tb'), ('x', 'p', 'cls_ivfs_tb')]


def _resp_gold():
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_resp.npz'))


def _close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)


def test_host_helpers_match_reference():
    """utils_spin.get_spin_matrix / spin_cls, utils.cl_inverse and qresp.get_qes against the unmodified reference
    (tests/golden/make_golden_resp.py), including TB / EB spectra (complex spin matrices)."""
    import golden_inputs as gi
    from plancklens_b200 import qresp, utils
    from plancklens_b200 import utils_spin as us
    g, r = _resp_gold(), gi.resp_case()
    for s1, s2 in [(0, 0), (0, 2), (2, 0), (2, 2), (2, -2), (-2, 2), (-2, 0), (0, -2)]:
        assert _close(us.get_spin_matrix(s1, s2, r['fal_tb']), g['spinmat_%d_%d' % (s1, s2)], 1e-15)
        assert _close(us.spin_cls(s1, s2, r['cls_ivfs_tb']), g['spincls_%d_%d' % (s1, s2)], 1e-15)
    inv = utils.cl_inverse(r['cls_dat'])
    assert set('clinv_' + k for k in inv) == set(k for k in g.files if k.startswith('clinv_'))
    for k, v in inv.items():
        assert _close(v, g['clinv_' + k], 1e-13)
    for key in ['ptt', 'p_p', 'p', 'x', 'ftt', 'pee', 'p_te', 'a_p']:
        qes = qresp.get_qes(key, r['lmax'], r['cls_weight'])
        assert len(qes) == int(g['qes_%s_n' % key][0]), key
        for i, q in enumerate(qes):
            assert [q.leg_a.spin_in, q.leg_a.spin_ou, q.leg_b.spin_in, q.leg_b.spin_ou] == list(g['qes_%s_%d_spins' % (key, i)])
            assert _close(q.leg_a.cl, g['qes_%s_%d_cla' % (key, i)], 1e-15) and _close(q.leg_b.cl, g['qes_%s_%d_clb' % (key, i)], 1e-15)
            assert _close(q.cL(np.arange(r['lmax_qlm'] + 1)), g['qes_%s_%d_cL' % (key, i)], 1e-15)


def _check_resp_and_nhl(tol):
    import golden_inputs as gi
    from plancklens_b200 import nhl, qresp
    g, r = _resp_gold(), gi.resp_case()
    for key, src, fal in RESP_KEYS:
        R = qresp.get_response(key, r['lmax'], src, r['cls_weight'], r['cls_len'], r[fal], lmax_qlm=r['lmax_qlm'])
        ref = g['resp_%s_%s_%s' % (key, src, fal)]
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(np.array(R) - ref)) <= tol * scale, (key, src, fal)
    for k1, k2, ivf in NHL_KEYS:
        N = nhl.get_nhl(k1, k2, r['cls_weight'], r[ivf], r['lmax'], r['lmax'], lmax_out=r['lmax_qlm'])
        ref = g['nhl_%s_%s_%s' % (k1, k2, ivf)]
        assert np.max(np.abs(np.array(N) - ref)) <= tol * np.max(np.abs(ref)), (k1, k2, ivf)


def test_response_and_n0_assembly_matches_reference_on_cpu(monkeypatch):
    """qresp.get_response / nhl.get_nhl host logic with the Wigner seam served by the CPU oracle on both sides: the
    port of the assembly code is checked without a GPU."""
    from oracle import ref_wigner
    from plancklens_b200 import utils_spin as us
    monkeypatch.setattr(us, 'wignerc', ref_wigner.wignerc)
    _check_resp_and_nhl(1e-12)


@pytest.mark.gpu
def test_response_and_n0_match_reference_on_gpu():
    """the same numbers with the Wigner transforms on the GPU"""
    _check_resp_and_nhl(1e-9)


def test_resp_and_nhl_libraries_cache_in_sqlite(tmp_path, monkeypatch):
    """resp_lib_simple / nhl_lib_simple: same numbers as the bare functions, cached in the reference's npdb layout
    (Wigner seam served by the oracle: runs on the CPU)."""
    import golden_inputs as gi
    from oracle import ref_wigner
    from plancklens_b200 import hp, nhl, qresp
    from plancklens_b200 import utils_spin as us
    calls = []

    def counted(*a, **k):
        calls.append(1)
        return ref_wigner.wignerc(*a, **k)
    monkeypatch.setattr(us, 'wignerc', counted)
    r = gi.resp_case()
    lib = qresp.resp_lib_simple(str(tmp_path / 'resp'), r['lmax'], r['cls_weight'], r['cls_len'], r['fal_sep'], r['lmax_qlm'])
    RG = lib.get_response('ptt', 'p')
    ref = qresp.get_response('ptt', r['lmax'], 'p', r['cls_weight'], r['cls_len'], r['fal_sep'], lmax_qlm=r['lmax_qlm'])[0]
    assert np.array_equal(RG, ref)
    n = len(calls)
    assert np.array_equal(lib.get_response('ptt', 'p'), ref) and len(calls) == n          # second call served by sqlite
    xc = lib.get_response('x_p', 'x')
    assert np.array_equal(xc, qresp.get_response('x_p', r['lmax'], 'x', r['cls_weight'], r['cls_len'], r['fal_sep'],
                                                 lmax_qlm=r['lmax_qlm'])[1])
    assert qresp.qe_spin_data('ptt')[:2] == (1, 'G') and qresp.qe_spin_data('x_p')[:2] == (1, 'C')
    assert qresp.qe_spin_data('ftt')[:2] == (0, 'G')
    # N0 from the empirical spectra of an in-memory filtering library
    q = gi.qe_case()
    iv = gi.idx_ivfs(q, hp)
    cls_w = gi.toy_cls(80)      # weights reach beyond lmax_ivf, as the CAMB tables the reference is used with
    nl = nhl.nhl_lib_simple(str(tmp_path / 'nhl'), iv, cls_w, 40)
    n0 = nl.get_sim_nhl(0, 'ptt', 'ptt')
    cls_emp = {'tt': hp.alm2cl(iv.get_sim_tlm(0))}
    ref = nhl.get_nhl('ptt', 'ptt', cls_w, cls_emp, len(cls_emp['tt']), len(cls_emp['tt']), lmax_out=40)[0]
    assert np.array_equal(n0, ref) and n0.shape == (41,) and np.all(n0[2:] > 0)
    n = len(calls)
    assert np.array_equal(nl.get_sim_nhl(0, 'ptt', 'ptt'), ref) and len(calls) == n


def _n0s_gold():
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_n0s.npz'))


def _check_n0s(tol, lmin=0):
    import golden_inputs as gi
    from plancklens_b200 import n0s
    g = _n0s_gold()
    expected = {'gmv': {'ptt', 'p_p', 'p'}, 'sep': {'ptt', 'p_p', 'p'}, 'tcut': {'ptt', 'p_p', 'p'}, 'curl': {'xtt', 'x_p', 'x'}}
    for name, kw in gi.n0s_cases().items():
        N0, N0c = n0s.get_N0(**kw)
        assert set(N0) == set(N0c) == expected[name]
        for k in N0:
            for tag, mine in (('G', N0[k]), ('C', N0c[k])):
                ref = g['%s_%s_%s' % (name, tag, k)]
                assert mine.shape == ref.shape
                assert np.max(np.abs(mine - ref)[lmin:]) <= tol * np.max(np.abs(ref)[lmin:]), (name, tag, k)


def test_n0s_match_reference_on_cpu(monkeypatch):
    """n0s.get_N0 (filters, filtered-map spectra with per-field multipole cuts, the Wiener-leg T cut, N0 = nhl / R^2)
    against the unmodified reference (tests/golden/make_golden_n0s.py), Wigner seam served by the oracle."""
    from oracle import ref_wigner
    from plancklens_b200 import utils_spin as us
    monkeypatch.setattr(us, 'wignerc', ref_wigner.wignerc)
    _check_n0s(1e-11)


@pytest.mark.gpu
def test_n0s_match_reference_on_gpu():
    """L >= 2: the L = 1 response of these band-limited toy cases is ~1e-6 of its L = 2 value, i.e. N0 = nhl / R^2
    amplifies the round-off of the Wigner transforms by ~1e11 there and the entry is numerical noise on both sides
    (the CPU test, where both sides share the oracle's transforms, does compare it)."""
    _check_n0s(1e-8, lmin=2)


def test_cls_dot_and_dls_conversions_match_reference():
    from plancklens_b200 import n0s, utils
    g = _n0s_gold()
    a = {'tt': np.arange(5.) + 1, 'ee': np.arange(7.) + 2, 'te': 0.1 * np.arange(6.), 'bb': 0.5 * np.ones(4)}
    b = {'tt': np.ones(7), 'ee': 2 * np.ones(7), 'bb': 3 * np.ones(7), 'tb': 0.2 * np.ones(7)}
    assert np.allclose(utils.cls_dot([a, b]), g['clsdot_ab'], rtol=1e-15, atol=0)
    assert np.allclose(utils.cls_dot([a, b, a]), g['clsdot_aba'], rtol=1e-14, atol=0)
    d = utils.cls_dot([a, b, a], ret_dict=True)
    assert set('clsdot_aba_' + k for k in d) == set(k for k in g.files if k.startswith('clsdot_aba_'))
    for k, v in d.items():
        assert np.allclose(v, g['clsdot_aba_' + k], rtol=1e-14, atol=0)
    dls, cldd = n0s.cls2dls({'tt': a['tt'], 'te': a['te'], 'pp': np.arange(8.) * 1e-3})
    assert np.allclose(dls, g['cls2dls'], rtol=1e-15, atol=0) and np.allclose(cldd, g['cls2dls_dd'], rtol=1e-15, atol=0)
    back = n0s.dls2cls(dls)
    assert np.allclose(back['tt'][1:5], a['tt'][1:5]) and np.allclose(back['te'][1:6], a['te'][1:6])
    with pytest.raises(NotImplementedError):
        n0s.get_N0_iter('p', 1., 1., 1., {}, 2, 10, 1)


def _check_resp2(tol):
    import golden_inputs as gi
    from plancklens_b200 import qresp
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_resp2.npz'))
    r = gi.resp_case()
    transf = gi.resp_transf(r['lmax'])

    def close(a, ref, what):
        assert np.max(np.abs(np.array(a) - ref)) <= tol * np.max(np.abs(ref)), what
    for key, src in gi.RESP2_CUSTOM:
        R = qresp.get_response(key, r['lmax'], src, r['cls_weight'], r['cls_len'], r['fal_sep'], lmax_qlm=r['lmax_qlm'], transf=transf)
        close(R, g['custom_%s_%s' % (key, src)], (key, src))
    for key, l, ck, src in gi.RESP2_DERIV:
        R = qresp.get_dresponse_dlncl(key, l, ck, r['lmax'], src, r['cls_weight'], r['cls_len'], r['fal_sep'], lmax_out=r['lmax_qlm'])
        close(R, g['dresp_%s_%d_%s_%s' % (key, l, ck, src)], (key, l, ck))
    for key in ('ptt', 'p_p'):
        GL, CL, terms = qresp.get_mf_resp(key, r['cls_len'], r['cls_ivfs_sep'], r['lmax'] - 10, r['lmax_qlm'], retterms=True)
        # the result is a difference of terms ~1e3 times larger: tolerance relative to the terms
        scale = np.max(np.abs(g['mf_%s_GK' % key]))
        assert np.max(np.abs(GL - g['mf_%s_G' % key])) <= tol * scale and np.max(np.abs(CL - g['mf_%s_C' % key])) <= tol * scale
        assert set(terms) == {'GK', 'GxiK', 'Gcons'}
        for k, v in terms.items():
            close(v, g['mf_%s_%s' % (key, k)], (key, k))
    assert len(qresp.get_mf_resp('ptt', r['cls_len'], r['cls_ivfs_sep'], 20, 30)) == 2
    with pytest.raises(AssertionError):
        qresp.get_mf_resp('p', r['cls_len'], r['cls_ivfs_sep'], 20, 30)


def test_custom_responses_derivatives_and_mf_response_on_cpu(monkeypatch):
    """qresp._get_response_custom ('n' source), get_dresponse_dlncl and get_mf_resp against the unmodified reference
    (tests/golden/make_golden_resp2.py), Wigner seam served by the oracle."""
    from oracle import ref_wigner
    from plancklens_b200 import utils_spin as us
    monkeypatch.setattr(us, 'wignerc', ref_wigner.wignerc)
    _check_resp2(1e-12)


@pytest.mark.gpu
def test_custom_responses_derivatives_and_mf_response_on_gpu():
    _check_resp2(1e-9)
