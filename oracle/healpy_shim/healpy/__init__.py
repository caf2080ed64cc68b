"""Stand-in for `healpy`, backed by the CPU oracle (TEST INFRASTRUCTURE ONLY).

Purpose: lets the UNMODIFIED reference Python under /root/reference be imported and run in the build container
(which has no healpy) to produce the golden vectors in tests/golden/ -- see tests/golden/make_golden.py.
Only the calls the reference hot path makes are provided.  Parity note: the transforms are the oracle's
(parity unpinned at that seam, see oracle/__init__.py); everything the reference computes ABOVE the seam is the
reference's own code.
"""
import numpy as np

from oracle import ref_geom as _rg
from oracle import ref_sht as _sht

from . import projector  # noqa: F401  (plancklens.utils imports healpy.projector.CartesianProj)



def alm2map(alms, nside, lmax=None, mmax=None, pol=True, **kw):
    """hp.alm2map; a (tlm, elm, blm) triple with pol=True gives (T, Q, U): a spin-0 and a spin-2 synthesis in the
    HEALPix polarization convention (the only multi-alm form the reference uses, qcinv/opfilt_tp.py:279)."""
    if isinstance(alms, (list, tuple)) or (isinstance(alms, np.ndarray) and alms.ndim == 2):
        assert len(alms) == 3 and pol
        L = _rg.alm_getlmax(np.asarray(alms[0]).size) if lmax is None else lmax
        q, u = _sht.alm2map_spin([alms[1], alms[2]], nside, 2, L)
        return [_sht.alm2map(alms[0], nside, lmax=L), q, u]
    return _sht.alm2map(alms, nside, lmax=lmax, mmax=mmax, **kw)


def map2alm(maps, lmax=None, mmax=None, iter=0, pol=True, **kw):
    """hp.map2alm(iter=0); (T, Q, U) with pol=True gives (tlm, elm, blm) (qcinv/opfilt_tp.py:22, :285)."""
    if isinstance(maps, (list, tuple)) or (isinstance(maps, np.ndarray) and maps.ndim == 2):
        assert len(maps) == 3 and pol and iter == 0
        e, b = _sht.map2alm_spin([maps[1], maps[2]], 2, lmax=lmax)
        return [_sht.map2alm(maps[0], lmax=lmax, iter=0), e, b]
    return _sht.map2alm(maps, lmax=lmax, mmax=mmax, iter=iter, **kw)


def smoothing(map_in, fwhm=0.0, sigma=None, iter=3, lmax=None, **kw):
    """hp.smoothing of a scalar map: map2alm with healpy's default of 3 refinement passes, Gaussian window, alm2map
    (plancklens/utils.py:296, :301)"""
    m = np.asarray(map_in, dtype=float)
    nside = _rg.npix2nside(m.size)
    lmax = 3 * nside - 1 if lmax is None else lmax
    if sigma is None:
        sigma = fwhm / np.sqrt(8.0 * np.log(2.0))
    ell = np.arange(lmax + 1)
    alm = _sht.map2alm(m, lmax=lmax, iter=iter)
    return _sht.alm2map(almxfl(alm, np.exp(-0.5 * ell * (ell + 1) * sigma ** 2)), nside, lmax=lmax)


alm2map_spin = _sht.alm2map_spin
map2alm_spin = _sht.map2alm_spin
nside2npix = _rg.nside2npix
npix2nside = _rg.npix2nside


def nside2pixarea(nside, degrees=False):
    a = 4 * np.pi / nside2npix(nside)
    return a * (180 / np.pi) ** 2 if degrees else a


class Alm:
    @staticmethod
    def getsize(lmax, mmax=None):
        return _rg.alm_getsize(lmax, mmax)

    @staticmethod
    def getlmax(s, mmax=None):
        return _rg.alm_getlmax(s)

    @staticmethod
    def getidx(lmax, l, m):
        return _rg.alm_getidx(lmax, l, m)


def _ls(lmax):
    return np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])


def almxfl(alm, fl, mmax=None, inplace=False):
    lmax = _rg.alm_getlmax(alm.size)
    f = np.zeros(lmax + 1, dtype=np.result_type(fl, float))
    n = min(lmax + 1, len(fl))
    f[:n] = np.asarray(fl)[:n]
    fac = f[_ls(lmax)]
    if inplace:
        alm *= fac
        return alm
    return alm * fac


def alm2cl(alms, alms2=None, lmax=None, mmax=None, lmax_out=None):
    a = np.asarray(alms)
    b = a if alms2 is None else np.asarray(alms2)
    L = _rg.alm_getlmax(a.size)
    w = np.full(a.size, 2.0)
    w[:L + 1] = 1.0
    cl = np.bincount(_ls(L), weights=w * (a * np.conj(b)).real, minlength=L + 1) / (2.0 * np.arange(L + 1) + 1)
    return cl if lmax_out is None else cl[:lmax_out + 1]


def gauss_beam(fwhm, lmax=512, pol=False):
    assert not pol
    sigma = fwhm / np.sqrt(8.0 * np.log(2.0))
    ell = np.arange(lmax + 1)
    return np.exp(-0.5 * ell * (ell + 1) * sigma ** 2)


def ud_grade(map_in, nside_out, power=None, **kw):
    assert power == -2, 'only the power=-2 (sum of children) form is used on the hot path'
    return _rg.ud_grade_sum(np.asarray(map_in, dtype=float), nside_out)


def write_alm(filename, alms, overwrite=True, **kw):
    with open(filename, 'wb') as f:
        np.save(f, np.asarray(alms))


def read_alm(filename, **kw):
    with open(filename, 'rb') as f:
        return np.load(f)


def write_map(filename, m, overwrite=True, **kw):
    with open(filename, 'wb') as f:
        np.save(f, np.asarray(m))


def read_map(filename, field=0, **kw):
    with open(filename, 'rb') as f:
        m = np.load(f)
    return m[field] if m.ndim == 2 else m
