"""Quadratic-estimator weight definitions (reference: plancklens/qresp.py:50-163).

`get_qes` and the response-leg helpers build the list of `utils_qe.qe` terms consumed by `utils_qe.qe_eval` (the
SHT hot path); `get_response` evaluates the estimator normalisations with the Wigner small-d transforms of
libplk_b200 (`utils_spin.wignerc`) where the reference calls its Fortran extension.
`resp_lib_simple` caches the same numbers in the reference's sqlite layout (`helpers/sql.py`).
"""
import numpy as np

import os
import pickle as pk

from . import utils as ut
from . import utils_qe as uqe
from . import utils_spin as uspin
from .helpers import mpi, sql


def _clinv(cl):
    cl = np.asarray(cl)
    ret = np.zeros_like(cl)
    nz = cl != 0
    ret[nz] = 1. / cl[nz]
    return ret


def get_resp_legs(source, lmax):
    r"""Spin response of the maps to an anisotropy source: for each input spin s in (0, -2, 2) the tuple
    (source spin r, response of +r, response of -r, scaling of G/C to the potential) (reference: qresp.py:94-121)."""
    if source in ['p', 'x']:
        # lensing: _sX -> _sX - 1/2 alpha_1 eth-bar _sX - 1/2 alpha_{-1} eth _sX
        return {s: (1, -0.5 * uspin.get_spin_lower(s, lmax), -0.5 * uspin.get_spin_raise(s, lmax),
                    lambda ell: uspin.get_spin_raise(0, np.max(ell))[ell]) for s in [0, -2, 2]}
    if source == 'f':
        half = 0.5 * np.ones(lmax + 1)
        return {s: (0, half.copy(), half.copy(), lambda ell: np.ones(len(ell))) for s in [0, -2, 2]}
    if source in ['a', 'a_p']:
        ret = {s: (0, -np.sign(s) * 1j * np.ones(lmax + 1), -np.sign(s) * 1j * np.ones(lmax + 1),
                   lambda ell: np.ones(len(ell))) for s in [-2, 2]}
        ret[0] = (0, np.zeros(lmax + 1), np.zeros(lmax + 1), lambda ell: np.ones(len(ell)))
        return ret
    assert 0, source + ' response legs not implemented'


def get_covresp(source, s1, s2, cls, lmax, transf=None):
    r"""Response of the spin covariance to an anisotropy source (reference: qresp.py:123-163)."""
    if source in ['p', 'x', 'f', 'a', 'a_p']:
        s_source, prR, mrR, cL_scal = get_resp_legs(source, lmax)[s1]
        coupl = uspin.spin_cls(s1, s2, cls)[:lmax + 1]
        return s_source, prR * coupl, mrR * coupl, cL_scal
    if source in ['stt', 's']:
        on = float(s1 == 0 and s2 == 0)
        quarter = 0.25 * on * np.ones(lmax + 1)
        return 0, quarter, quarter.copy(), lambda ell: np.ones(len(ell))
    assert 0, 'source ' + source + ' cov. response not implemented'


def get_qes(qe_key, lmax, cls_weight, lmax2=None, transf=None):
    """List of `utils_qe.qe` terms defining the estimator `qe_key` (e.g. 'ptt', 'p_p', 'p', 'pee', 'p_eb', ...).

    The weights act on the inverse-variance filtered spin-weight maps (reference: qresp.py:50-92)."""
    if lmax2 is None:
        lmax2 = lmax
    if qe_key[0] in ['p', 'x', 'a', 'f', 's']:
        if qe_key in ['ptt', 'xtt', 'att', 'ftt', 'stt']:
            s_lefts = [0]
        elif qe_key in ['p_p', 'x_p', 'a_p', 'f_p']:
            s_lefts = [-2, 2]
        else:
            s_lefts = [0, -2, 2]
        qes = []
        for s_left in s_lefts:
            for sin in s_lefts:
                sout = -s_left
                s_qe, _, cl_sosi, cL_out = get_covresp(qe_key[0], sout, sin, cls_weight, lmax2, transf=transf)
                if np.any(cl_sosi):
                    lega = uqe.qeleg(s_left, s_left, 0.5 * (1. + (s_left == 0)) * np.ones(lmax + 1))
                    legb = uqe.qeleg(sin, sout + s_qe, 0.5 * (1. + (sin == 0)) * 2 * cl_sosi)
                    qes.append(uqe.qe(lega, legb, cL_out))
        rest = qe_key[1:]
        if len(qe_key) == 1 or rest in ['tt', '_p']:
            return uqe.qe_simplify(qes)
        if rest in ['te', 'et', 'tb', 'bt', 'ee', 'eb', 'be', 'bb']:
            return uqe.qe_simplify(uqe.qe_proj(qes, qe_key[1], qe_key[2]))
        if rest in ['_te', '_tb', '_eb']:
            return uqe.qe_simplify(uqe.qe_proj(qes, qe_key[2], qe_key[3]) + uqe.qe_proj(qes, qe_key[3], qe_key[2]))
        assert 0, 'qe key %s  not recognized' % qe_key
    if qe_key == 'ntt':
        lega = uqe.qeleg(0, 0, 1 * _clinv(transf[:lmax + 1]))
        legb = uqe.qeleg(0, 0, 0.5 * _clinv(transf[:lmax + 1]))
        return uqe.qe_simplify([uqe.qe(lega, legb, lambda L: np.ones(len(L)))])
    assert 0, qe_key + ' not implemented'


def get_response(qe_key, lmax_ivf, source, cls_weight, cls_cmb, fal, fal_leg2=None, lmax_ivf2=None, lmax_qlm=None,
                 transf=None):
    r"""QE response (normalisation) :math:`R_L` of estimator `qe_key` to anisotropy `source` (reference:
    qresp.py:269-311).  Returns (GG, CC, GC, CG).  Not symmetrised in the two filters if they differ."""
    if lmax_ivf2 is None:
        lmax_ivf2 = lmax_ivf
    if lmax_qlm is None:
        lmax_qlm = lmax_ivf + lmax_ivf2
    if '_bh_' in qe_key:     # bias-hardened estimators (reference: qresp.py:289-307)
        k, hsource = qe_key.split('_bh_')
        assert len(hsource) == 1, hsource
        h = hsource[0]
        kw = dict(fal_leg2=fal_leg2, lmax_ivf2=lmax_ivf2, lmax_qlm=lmax_qlm, transf=transf)
        GG_ks, CC_ks, GC_ks, CG_ks = get_response(k, lmax_ivf, source, cls_weight, cls_cmb, fal, **kw)
        GG_hs, CC_hs, GC_hs, CG_hs = get_response(h + k[1:], lmax_ivf, source, cls_weight, cls_cmb, fal, **kw)
        GG_kh, CC_kh, GC_kh, CG_kh = get_response(k, lmax_ivf, h, cls_weight, cls_cmb, fal, **kw)
        GG_hh, CC_hh, GC_hh, CG_hh = get_response(h + k[1:], lmax_ivf, h, cls_weight, cls_cmb, fal, **kw)
        iG, iC = ut.cli(GG_hh), ut.cli(CC_hh)
        return (GG_ks - (GG_kh * GG_hs * iG + GC_kh * CG_hs * iC), CC_ks - (CG_kh * GC_hs * iG + CC_kh * CC_hs * iC),
                GC_ks - (GG_kh * GC_hs * iG + GC_kh * CC_hs * iC), CG_ks - (CG_kh * GG_hs * iG + CC_kh * CG_hs * iC))
    qes = get_qes(qe_key, lmax_ivf, cls_weight, lmax2=lmax_ivf2, transf=transf)
    custom = _get_response_custom(qe_key, qes, source, fal, lmax_qlm, fal_leg2=fal_leg2, transf=transf)
    if custom is not None:
        return custom
    return _get_response(qes, source, cls_cmb, fal, lmax_qlm, fal_leg2=fal_leg2)


def _get_response_custom(qe_key, qes, source, fal_leg1, lmax_qlm, fal_leg2=None, transf=None):
    """Responses that do not fit the (source spin, covariance response) scheme of `get_covresp`: temperature
    estimators responding to a noise-variance map ('n' / 'ntt'), a spin-0 source acting on the beam-deconvolved
    temperature (reference: qresp.py:315-361).  None for every other combination."""
    if not ('tt' in qe_key and source in ['n', 'ntt']):
        return None
    assert transf is not None
    fal_leg2 = fal_leg1 if fal_leg2 is None else fal_leg2
    Ls = np.arange(lmax_qlm + 1, dtype=int)
    R = np.zeros((4, lmax_qlm + 1), dtype=float)          # GG, CC, GC, CG
    bi = _clinv(transf)
    for qe in qes:
        si, ti, so, to = qe.leg_a.spin_in, qe.leg_b.spin_in, qe.leg_a.spin_ou, qe.leg_b.spin_ou
        assert (si, ti) == (0, 0)
        spin_qe = abs(so + to)

        def term(sgn):
            FA = uspin.get_spin_matrix(sgn * si, 0, fal_leg1)
            FB = uspin.get_spin_matrix(sgn * ti, 0, fal_leg2)
            cla, clb = (qe.leg_a.cl, qe.leg_b.cl) if sgn > 0 else (qe.leg_a.cl.conj(), qe.leg_b.cl.conj())
            return FB, uspin.wignerc(ut.joincls([cla, FA, bi]), ut.joincls([clb, FB, bi]), sgn * so, 0, sgn * to, 0,
                                     lmax_out=lmax_qlm)
        FB, Rp = term(+1)
        if not np.any(FB):
            continue
        Rm = (-1) ** (so + si + to + ti) * term(-1)[1] if spin_qe > 0 else Rp
        w, sg = 0.5 * qe.cL(Ls), (-1) ** spin_qe
        R[0] += w * (Rp.real + sg * Rm.real)
        R[1] += w * (Rp.real - sg * Rm.real)
        R[2] += w * (-Rp.imag + sg * Rm.imag)
        R[3] += w * (Rp.imag + sg * Rm.imag)
    return R[0], R[1], R[2], R[3]


def get_dresponse_dlncl(qe_key, l, cl_key, lmax_ivf, source, cls_weight, cls_cmb, fal_leg1, fal_leg2=None, lmax_ivf2=None,
                        lmax_out=None):
    r"""Derivative :math:`dR_L / d\ln C_\ell` of the isotropic response with respect to one multipole of one CMB
    spectrum (reference: qresp.py:364-374): the response to a spectrum that is zero except for `cls_cmb[cl_key][l]`."""
    if lmax_ivf2 is None:
        lmax_ivf2 = lmax_ivf
    if lmax_out is None:
        lmax_out = lmax_ivf + lmax_ivf2
    dcls = {k: np.zeros_like(cls_cmb[k]) for k in cls_cmb}
    dcls[cl_key][l] = cls_cmb[cl_key][l]
    qes = get_qes(qe_key, lmax_ivf, cls_weight, lmax2=lmax_ivf2)
    return _get_response(qes, source, dcls, fal_leg1, lmax_out, fal_leg2=fal_leg2)


def get_mf_resp(qe_key, cls_cmb, cls_ivfs, lmax_qe, lmax_out, retterms=False):
    """Deflection-induced mean-field response of the lensing estimators 'ptt' and 'p_p' (reference: qresp.py:421-500).

    Two groups of Wigner correlation functions: (xi K xi - xi)(K)-like terms, built from the filtered-map spectra and
    the CMB spectra minus their filtered part, and (xi K)(xi K)-like terms; the L = 1 value of the curl part, which must
    vanish, is subtracted from both.  Returns (gradient, curl) responses, with the three terms if `retterms`."""
    assert qe_key in ['p_p', 'ptt'], qe_key
    if qe_key == 'ptt':
        lmax_cmb, spins, keys = len(cls_cmb['tt']) - 1, [0], ['tt']
    else:
        lmax_cmb, spins, keys = min(len(cls_cmb['ee']) - 1, len(cls_cmb['bb'] - 1)), [-2, 2], ['ee', 'bb']
    assert lmax_qe <= lmax_cmb
    n = lmax_qe + 1
    cl_cic = {k: cls_cmb[k][:n] ** 2 * cls_ivfs[k][:n] for k in keys}      # C F C
    cl_ci = {k: cls_cmb[k][:n] * cls_ivfs[k][:n] for k in keys}            # C F
    half = lambda sp: 0.5 if sp != 0 else 1.0     # each B in B Cov^-1 B^dagger maps spin fields to T E B with a factor 1/2

    def ladder(a, sp, lmax):
        return uspin.get_spin_lower(sp, lmax) if a == -1 else uspin.get_spin_raise(sp, lmax)

    G1, C1 = np.zeros(lmax_out + 1), np.zeros(lmax_out + 1)
    G2, C2 = np.zeros(lmax_out + 1), np.zeros(lmax_out + 1)
    for s1 in spins:
        for s2 in spins:
            sgn = 2 * (-1) ** (s1 + s2)
            # (xi K xi - xi)(K)
            cl1 = uspin.spin_cls(s1, s2, cls_ivfs)[:n] * (half(s1) * half(s2))
            cl2 = np.copy(uspin.spin_cls(s2, s1, cls_cmb)[:lmax_cmb + 1])
            cl2[:n] -= uspin.spin_cls(s2, s1, cl_cic)[:n]       # subtracted before the transform: unstable otherwise
            if np.any(cl1) and np.any(cl2):
                aj = ladder(-1, -s1, lmax_cmb)
                for a in (-1, 1):                               # the b = -1 terms follow from the a <-> b symmetry
                    hL = sgn * uspin.wignerc(cl1, cl2 * ladder(a, s2, lmax_cmb) * aj, s2, s1, -s2 - a, -s1 - 1, lmax_out=lmax_out)
                    G1 -= a * hL
                    C1 -= hL
            # (xi K)(xi K)
            cl1 = uspin.spin_cls(s2, s1, cl_ci)[:n] * half(s1)
            cl2 = uspin.spin_cls(s1, s2, cl_ci)[:n] * half(s2)
            if np.any(cl1) and np.any(cl2):
                aj = ladder(-1, s1, lmax_qe)
                for a in (-1, 1):
                    hL = sgn * uspin.wignerc(cl1 * ladder(a, s2, lmax_qe), cl2 * aj, -s2 - a, -s1, s2, s1 - 1, lmax_out=lmax_out)
                    G2 -= a * hL
                    C2 -= hL
    terms = {'GK': G1.copy(), 'GxiK': -G2}
    GL, CL = G1 - G2, C1 - C2
    terms['Gcons'] = -np.ones_like(GL) * CL[1]
    print("CL[1] ", CL[1])
    print("GL[1] (before subtraction) ", GL[1])
    print("GL[1] (after subtraction) ", GL[1] - CL[1])
    GL = GL - CL[1]
    CL = CL - CL[1]
    fac = 0.25 * np.arange(lmax_out + 1) * np.arange(1, lmax_out + 2)
    GL, CL = GL * fac, CL * fac
    for k in terms:
        terms[k] = terms[k] * fac
    return (GL, CL, terms) if retterms else (GL, CL)


def _get_response(qes, source, cls_cmb, fal_leg1, lmax_qlm, fal_leg2=None):
    """reference: qresp.py:376-417 (same loops, same order of accumulation)"""
    fal_leg2 = fal_leg1 if fal_leg2 is None else fal_leg2
    RGG = np.zeros(lmax_qlm + 1, dtype=float)
    RCC = np.zeros(lmax_qlm + 1, dtype=float)
    RGC = np.zeros(lmax_qlm + 1, dtype=float)
    RCG = np.zeros(lmax_qlm + 1, dtype=float)
    Ls = np.arange(lmax_qlm + 1, dtype=int)
    for qe in qes:
        si, ti = (qe.leg_a.spin_in, qe.leg_b.spin_in)
        so, to = (qe.leg_a.spin_ou, qe.leg_b.spin_ou)
        for s2 in [0, -2, 2]:
            FA = uspin.get_spin_matrix(si, s2, fal_leg1)
            if not np.any(FA):
                continue
            for t2 in [0, -2, 2]:
                FB = uspin.get_spin_matrix(ti, t2, fal_leg2)
                if not np.any(FB):
                    continue
                rW_st, prW_st, mrW_st, s_cL_st = get_covresp(source, -s2, t2, cls_cmb, len(FB) - 1)
                clA = ut.joincls([qe.leg_a.cl, FA])
                clB = ut.joincls([qe.leg_b.cl, FB, mrW_st.conj()])
                Rpr_st = uspin.wignerc(clA, clB, so, s2, to, -s2 + rW_st, lmax_out=lmax_qlm) * s_cL_st(Ls)

                rW_ts, prW_ts, mrW_ts, s_cL_ts = get_covresp(source, -t2, s2, cls_cmb, len(FA) - 1)
                clA = ut.joincls([qe.leg_a.cl, FA, mrW_ts.conj()])
                clB = ut.joincls([qe.leg_b.cl, FB])
                Rpr_st = Rpr_st + uspin.wignerc(clA, clB, so, -t2 + rW_ts, to, t2, lmax_out=lmax_qlm) * s_cL_ts(Ls)
                assert rW_st == rW_ts and rW_st >= 0, (rW_st, rW_ts)
                if rW_st > 0:
                    clA = ut.joincls([qe.leg_a.cl, FA])
                    clB = ut.joincls([qe.leg_b.cl, FB, prW_st.conj()])
                    Rmr_st = uspin.wignerc(clA, clB, so, s2, to, -s2 - rW_st, lmax_out=lmax_qlm) * s_cL_st(Ls)
                    clA = ut.joincls([qe.leg_a.cl, FA, prW_ts.conj()])
                    clB = ut.joincls([qe.leg_b.cl, FB])
                    Rmr_st = Rmr_st + uspin.wignerc(clA, clB, so, -t2 - rW_ts, to, t2, lmax_out=lmax_qlm) * s_cL_ts(Ls)
                else:
                    Rmr_st = Rpr_st
                prefac = qe.cL(Ls)
                RGG += prefac * (Rpr_st.real + Rmr_st.real * (-1) ** rW_st)
                RCC += prefac * (Rpr_st.real - Rmr_st.real * (-1) ** rW_st)
                RGC += prefac * (-Rpr_st.imag + Rmr_st.imag * (-1) ** rW_st)
                RCG += prefac * (Rpr_st.imag + Rmr_st.imag * (-1) ** rW_st)
    return RGG, RCC, RGC, RCG


def qe_spin_data(qe_key):
    """(spin of the estimator output, 'G' or 'C', unique input spins >= 0, spin-1 key) (reference: qresp.py:164-179)."""
    if qe_key in ['ntt']:
        return 0, 'G', [0], 'n'
    qes = get_qes(qe_key, 10, {k: np.ones(11 + 4, dtype=float) for k in ['tt', 'te', 'ee', 'bb']})
    spins_out = [qe.leg_a.spin_ou + qe.leg_b.spin_ou for qe in qes]
    spins_in = np.unique(np.abs([qe.leg_a.spin_in for qe in qes] + [qe.leg_b.spin_in for qe in qes]))
    assert len(np.unique(spins_out)) == 1, spins_out
    assert spins_out[0] >= 0, spins_out[0]
    if spins_out[0] > 0:
        assert qe_key[0] in ['x', 'p'], 'non-zero spin anisotropy ' + qe_key + ' not implemented ?'
    return spins_out[0], 'C' if qe_key[0] == 'x' else 'G', spins_in, 'p' if qe_key[0] == 'x' else qe_key[0]


class resp_lib_simple:
    """Cached QE responses (reference: qresp.py:182-266): wraps `get_response`, results in `npdb.db`."""

    def __init__(self, lib_dir, lmax_ivf, cls_weight, cls_cmb, fal, lmax_qlm, transf=None):
        self.lmax_qe = lmax_ivf
        self.lmax_qlm = lmax_qlm
        self.cls_weight = cls_weight
        self.cls_cmb = cls_cmb
        self.fal = fal
        self.transf = transf
        self.lib_dir = lib_dir
        fn_hash = os.path.join(lib_dir, 'resp_hash.pk')
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            if not os.path.exists(fn_hash):
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fn_hash, 'rb') as f:
            ut.hash_check(pk.load(f), self.hashdict(), fn=fn_hash)
        self.npdb = sql.npdb(os.path.join(lib_dir, 'npdb.db'))

    def hashdict(self):
        ret = {'lmaxqe': self.lmax_qe, 'lmax_qlm': self.lmax_qlm}
        for k in self.cls_weight.keys():
            ret['clsweight ' + k] = ut.clhash(self.cls_weight[k])
        for k in self.cls_cmb.keys():
            ret['clscmb ' + k] = ut.clhash(self.cls_cmb[k])
        for k in self.fal.keys():
            ret['fal' + k] = ut.clhash(self.fal[k])
        return ret

    def get_response(self, k, ksource, recache=False):
        """Response of estimator key k to anisotropy source ksource (GG part for gradient keys, CC for curl keys)."""
        if '_bh_' in k:      # bias-hardened estimator
            kQE, bhksource = k.split('_bh_')
            assert len(ksource) == 1, (kQE, ksource)
            wL = self.get_response(kQE, bhksource, recache=recache)
            wL = wL * ut.cli(self.get_response(bhksource + kQE[1:], bhksource, recache=recache))
            ret = np.copy(self.get_response(kQE, ksource, recache=recache))
            ret -= wL * self.get_response(bhksource + kQE[1:], ksource, recache=recache)
            return ret
        if k in ['xmtt', 'pmtt']:
            return self.get_response(k[0], ksource, recache=recache) - self.get_response(k[0] + 'tt', ksource, recache=recache)
        s, GorC, sins, ksp = qe_spin_data(k)
        assert s >= 0, s
        if s == 0:
            assert GorC == 'G', (s, GorC)
        base = 'qe_' + ksp + k[1:] + '_source_%s' % ksource
        fn = base + '_' + GorC + GorC
        if self.npdb.get(fn) is None or recache:
            GG, CC, GC, CG = get_response(k, self.lmax_qe, ksource, self.cls_weight, self.cls_cmb, self.fal,
                                          lmax_qlm=self.lmax_qlm, transf=self.transf)
            if np.any(CG) or np.any(GC):
                print("Warning: C-G or G-C responses non-zero but not returned")
            if recache and self.npdb.get(fn) is not None:
                self.npdb.remove(base + '_GG')
                if s > 0:
                    self.npdb.remove(base + '_CC')
            self.npdb.add(base + '_GG', GG)
            if s > 0:
                self.npdb.add(base + '_CC', CC)
        return self.npdb.get(fn)
