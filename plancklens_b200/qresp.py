"""Quadratic-estimator weight definitions (reference: plancklens/qresp.py:50-163).

Only the part of `qresp` that is on the SHT hot path is mirrored: `get_qes` and the response-leg helpers that
build the list of `utils_qe.qe` terms consumed by `utils_qe.qe_eval`.  The response / normalisation integrals
(`get_response`, `resp_lib_simple`) need the Wigner small-d transforms and are out of scope.
"""
import numpy as np

from . import utils_qe as uqe
from . import utils_spin as uspin


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
