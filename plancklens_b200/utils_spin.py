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
