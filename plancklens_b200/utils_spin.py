r"""Spin-weight utilities (reference: plancklens/utils_spin.py).

Conventions, as in the reference: :math:`_{\pm |s|} X_{lm} = - (\pm)^{|s|} (G_{lm} \pm i C_{lm})`, hence
:math:`G^{0}_{lm} = -T_{lm}`, :math:`G^{2}_{lm} = E_{lm}`, :math:`C^{2}_{lm} = B_{lm}`.
"""
import numpy as np

from . import hp

HASWIGNER = False   # Wigner small-d integrals (qresp / nhl) are outside the hot path


def alm2map_spin(gclm, nside, spin, lmax, mmax=None):
    """Spin >= 0 synthesis; spin 0 follows G^0 = -T (reference: utils_spin.py:21-27)."""
    assert spin >= 0, spin
    assert len(gclm) == 2, len(gclm)
    if spin > 0:
        return hp.alm2map_spin(gclm, nside, spin, lmax, mmax=mmax)
    return hp.alm2map(-np.asarray(gclm[0]), nside, lmax=lmax, mmax=mmax), 0.


def map2alm_spin(maps, spin, lmax=None, mmax=None):
    """Spin >= 0 analysis (reference: utils_spin.py:29-34)."""
    assert spin >= 0, spin
    if spin > 0:
        return hp.map2alm_spin(maps, spin, lmax=lmax, mmax=mmax)
    return -hp.map2alm(maps[0], lmax=lmax, mmax=mmax, iter=0), 0.


def get_spin_raise(s, lmax):
    r""":math:`\sqrt{(l - s)(l + s + 1)}` for |s| <= l <= lmax (reference: utils_spin.py:96)."""
    ret = np.zeros(lmax + 1)
    l = np.arange(abs(s), lmax + 1, dtype=float)
    ret[abs(s):] = np.sqrt((l - s) * (l + s + 1))
    return ret


def get_spin_lower(s, lmax):
    r""":math:`-\sqrt{(l + s)(l - s + 1)}` for |s| <= l <= lmax (reference: utils_spin.py:106)."""
    ret = np.zeros(lmax + 1)
    l = np.arange(abs(s), lmax + 1, dtype=float)
    ret[abs(s):] = -np.sqrt((l + s) * (l - s + 1))
    return ret


def _transposed(cls):
    return {(k + k if len(k) == 1 else k[1] + k[0]): np.copy(v) for k, v in cls.items()}


def spin_cls(s1, s2, cls):
    r"""Spin-weighted spectrum :math:`_{s1}X_{lm}\, _{s2}X^{*}_{lm}` from T, E, B spectra (reference: utils_spin.py:127)."""
    if s1 < 0:
        return (-1) ** (s1 + s2) * np.conjugate(spin_cls(-s1, -s2, _transposed(cls)))
    assert s1 in (0, 2) and s2 in (0, -2, 2), (s1, s2, 'not implemented')
    if s1 == 0:
        if s2 == 0:
            return cls['tt']
        te = cls['te'] if 'te' in cls else cls['et']
        tb = cls.get('tb')
        return -te if tb is None else -te + 1j * np.sign(s2) * tb
    if s2 == 0:
        et = cls['et'] if 'et' in cls else cls['te']
        tb = cls.get('bt', cls.get('tb'))
        return -et if tb is None else -et - 1j * tb
    if s2 == 2:
        return cls['ee'] + cls['bb']
    eb = cls.get('be', cls.get('eb'))
    return cls['ee'] - cls['bb'] if eb is None else cls['ee'] - cls['bb'] + 2j * eb


def get_spin_matrix(sout, sin, cls):
    r"""Spin-space matrix R^{-1} cls[T, E, B] R, R mapping _{0, \pm 2}X to T, E, B (reference: utils_spin.py:158-197).
    Missing spectra count as zero; 't', 'e', 'b' keys are accepted for 'tt', 'ee', 'bb'."""
    assert sin in [0, 2, -2] and sout in [0, 2, -2], (sin, sout)
    ee = cls.get('ee', cls.get('e', 0.))
    bb = cls.get('bb', cls.get('b', 0.))
    te = cls.get('te', 0.)
    tb, eb = cls.get('tb', None), cls.get('eb', None)
    if sin == 0:
        if sout == 0:
            return cls.get('tt', cls.get('t', 0.))
        return (-te - 1j * np.sign(sout) * tb) if tb is not None else -te
    if sout == 0:
        if tb is None:
            return -0.5 * te
        return -0.5 * (te - 1j * tb) if sin == 2 else -0.5 * (te + 1j * tb)
    if sout == sin:
        return 0.5 * (ee + bb)
    ret = 0.5 * (ee - bb)
    if eb is None:
        return ret
    return ret - 1j * eb if sin == 2 else ret + 1j * eb


_GL_cache = {}


def wignerc(cl1, cl2, sp1, s1, sp2, s2, lmax_out=None):
    r"""Legendre coefficients of :math:`(\xi_{sp1,s1} \xi_{sp2,s2})(\cos\theta)` from their harmonic series, exact by
    Gauss-Legendre quadrature (reference: utils_spin.py:52-93).  The three Wigner transforms run on the GPU
    (`plancklens_b200.wigners`, libplk_b200) instead of the reference's Fortran extension."""
    from . import wigners
    cl1, cl2 = np.asarray(cl1), np.asarray(cl2)
    lmax1, lmax2 = len(cl1) - 1, len(cl2) - 1
    lmax_out = lmax1 + lmax2 if lmax_out is None else lmax_out
    lmaxtot = lmax1 + lmax2 + lmax_out
    if not (np.any(cl1) and np.any(cl2)):
        return np.zeros(lmax_out + 1, dtype=float)
    N = (lmaxtot + 2 - lmaxtot % 2) // 2
    if N not in _GL_cache:
        import torch
        xg, wg = wigners.get_xgwg(-1., 1., N)
        _GL_cache[N] = (xg, wg, torch.from_numpy(xg).cuda())
    xg, wg, xd = _GL_cache[N]

    def pos(cl, a, b):
        if np.iscomplexobj(cl):
            return wigners.wignerpos(np.real(cl), xd, a, b) + 1j * wigners.wignerpos(np.imag(cl), xd, a, b)
        return wigners.wignerpos(cl, xd, a, b)
    xi1xi2w = pos(cl1, sp1, s1) * pos(cl2, sp2, s2) * wg
    spo, so = sp1 + sp2, s1 + s2
    if np.iscomplexobj(xi1xi2w):
        return wigners.wignercoeff(np.real(xi1xi2w), xd, spo, so, lmax_out) \
            + 1j * wigners.wignercoeff(np.imag(xi1xi2w), xd, spo, so, lmax_out)
    return wigners.wignercoeff(xi1xi2w, xd, spo, so, lmax_out)
