"""The slice of the healpy surface the plancklens hot path touches, without healpy.

Transforms (`alm2map`, `map2alm`, `alm2map_spin`, `map2alm_spin`) run on the GPU through libplk_b200;
the small helpers (`Alm`, `almxfl`, `alm2cl`, `gauss_beam`, `nside2npix`, `ud_grade`, ...) are host numpy.
File I/O uses numpy's .npy container under whatever file name the caller passes (FITS is out of scope,
SURVEY.md section 8f rank 4).
"""
import numpy as np

UNSEEN = -1.6375e30


# ------------------------------------------------------------------ sizes
def nside2npix(nside):
    return 12 * int(nside) ** 2


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    return nside


def nside2pixarea(nside, degrees=False):
    a = 4.0 * np.pi / nside2npix(nside)
    return a * (180.0 / np.pi) ** 2 if degrees else a


def nside2resol(nside, arcmin=False):
    r = np.sqrt(nside2pixarea(nside))
    return np.rad2deg(r) * 60.0 if arcmin else r


class Alm:
    """Index helpers for the m-major triangular alm layout (only m >= 0 stored)."""

    @staticmethod
    def getsize(lmax, mmax=None):
        mmax = lmax if mmax is None or mmax < 0 or mmax > lmax else mmax
        return mmax * (2 * lmax + 1 - mmax) // 2 + lmax + 1

    @staticmethod
    def getlmax(s, mmax=None):
        if mmax is not None and mmax >= 0:
            x = (2 * s + mmax ** 2 - mmax - 2) / (2 * mmax + 2)
        else:
            x = (-3 + np.sqrt(1 + 8 * s)) / 2
        return int(x) if x == np.floor(x) else -1

    @staticmethod
    def getidx(lmax, l, m):
        return m * (2 * lmax + 1 - m) // 2 + l

    @staticmethod
    def getlm(lmax, i=None):
        sz = Alm.getsize(lmax)
        i = np.arange(sz) if i is None else np.asarray(i)
        m = (np.ceil(((2 * lmax + 1) - np.sqrt((2 * lmax + 1) ** 2 - 8 * (i - lmax))) / 2)).astype(int)
        l = i - m * (2 * lmax + 1 - m) // 2
        return l, m


def _ls_of(lmax):
    return np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])


_LS_CACHE = {}


def alm_ls(lmax):
    if lmax not in _LS_CACHE:
        if len(_LS_CACHE) > 8:
            _LS_CACHE.clear()
        _LS_CACHE[lmax] = _ls_of(lmax)
    return _LS_CACHE[lmax]


def almxfl(alm, fl, mmax=None, inplace=False):
    """alm_lm * fl_l.  A short fl multiplies the higher l by zero (healpy behaviour)."""
    alm = np.asarray(alm) if not inplace else alm
    lmax = Alm.getlmax(alm.size, mmax)
    assert lmax >= 0, 'alm size does not match a triangular layout'
    fl = np.asarray(fl)
    f = np.zeros(lmax + 1, dtype=fl.dtype if np.iscomplexobj(fl) else float)
    n = min(lmax + 1, fl.size)
    f[:n] = fl[:n]
    fac = f[alm_ls(lmax)]
    if inplace:
        alm *= fac
        return alm
    return alm * fac


def alm2cl(alms, alms2=None, lmax=None, mmax=None, lmax_out=None):
    """(Cross-)spectrum C_l = 1/(2l+1) sum_m a_lm conj(b_lm) for real fields."""
    a = np.asarray(alms)
    b = a if alms2 is None else np.asarray(alms2)
    assert a.size == b.size
    L = Alm.getlmax(a.size) if lmax is None else lmax
    lo = L if lmax_out is None else lmax_out
    ls = alm_ls(L)
    w = np.full(a.size, 2.0)
    w[:L + 1] = 1.0
    prod = w * (a * np.conj(b)).real
    cl = np.bincount(ls, weights=prod, minlength=L + 1) / (2.0 * np.arange(L + 1) + 1.0)
    return cl[:lo + 1]


def gauss_beam(fwhm, lmax=512, pol=False):
    sigma = fwhm / np.sqrt(8.0 * np.log(2.0))
    ell = np.arange(lmax + 1)
    g = np.exp(-0.5 * ell * (ell + 1) * sigma ** 2)
    if not pol:
        return g
    f = np.exp([0.0, 2 * sigma ** 2, 2 * sigma ** 2, sigma ** 2])
    return g[:, None] * f[None, :]


def pixwin(nside, pol=False, lmax=None):
    raise NotImplementedError("hp.pixwin needs the HEALPix data files, which are not available here; use a "
                              "beam-only transfer function (SURVEY.md section 7.3 item 6)")


# ------------------------------------------------------------------ pixel orderings (setup only)
def _interleave0(v):
    v = v.astype(np.int64)
    for s, msk in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                   (2, 0x3333333333333333), (1, 0x5555555555555555)):
        v = (v | (v << s)) & msk
    return v


def ring2nest(nside, ipix):
    """RING -> NEST pixel index (nside a power of two), vectorised."""
    N = int(nside)
    p = np.asarray(ipix, dtype=np.int64)
    npix, ncap, n4 = 12 * N * N, 2 * N * (N - 1), 4 * N
    ring = np.empty_like(p); iphi = np.empty_like(p); nr = np.empty_like(p)
    ksh = np.zeros_like(p); face = np.empty_like(p)
    cn, cs = p < ncap, p >= npix - ncap
    ce = ~(cn | cs)
    if cn.any():
        q = p[cn]
        r = (1 + np.sqrt(1 + 2 * q.astype(float)).astype(np.int64)) // 2
        r = np.where(2 * r * (r - 1) > q, r - 1, r)
        r = np.where(2 * r * (r + 1) <= q, r + 1, r)
        ring[cn] = r; nr[cn] = r
        iphi[cn] = q + 1 - 2 * r * (r - 1)
        face[cn] = (iphi[cn] - 1) // r
    if ce.any():
        q = p[ce] - ncap
        t = q // n4
        r = t + N
        ring[ce] = r; nr[ce] = N
        ph = q - t * n4 + 1
        iphi[ce] = ph
        ksh[ce] = (r + N) & 1
        ire, irm = t + 1, 2 * N + 1 - t
        ifm = (ph - ire // 2 + N - 1) // N
        ifp = (ph - irm // 2 + N - 1) // N
        face[ce] = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))
    if cs.any():
        q = npix - p[cs]
        r = (1 + np.sqrt(2 * q.astype(float) - 1).astype(np.int64)) // 2
        r = np.where(2 * r * (r - 1) >= q, r - 1, r)
        r = np.where(2 * r * (r + 1) < q, r + 1, r)
        nr[cs] = r
        iphi[cs] = 4 * r + 1 - (q - 2 * r * (r - 1))
        face[cs] = 8 + (iphi[cs] - 1) // r
        ring[cs] = 4 * N - r
    jr = 2 + (face >> 2)
    jp = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])[face]
    irt = ring - jr * N + 1
    ipt = 2 * iphi - jp * nr - ksh - 1
    ipt = np.where(ipt >= 2 * N, ipt - 8 * N, ipt)
    ix = (ipt - irt) >> 1
    iy = (-(ipt + irt)) >> 1
    return face * N * N + _interleave0(ix) + (_interleave0(iy) << 1)


def ud_grade(map_in, nside_out, pess=False, order_in='RING', order_out=None, power=None, dtype=None):
    """Degrade (or keep) a RING map.  power=-2 sums the children (reference: opfilt_tt.py:179, opfilt_pp.py:251);
    power=None averages them."""
    assert order_in == 'RING' and order_out in (None, 'RING')
    m = np.asarray(map_in, dtype=float)
    nside_in = npix2nside(m.size)
    if nside_out == nside_in:
        return m.copy()
    if nside_out > nside_in:
        raise NotImplementedError("ud_grade: upgrading is not on the hot path")
    nest = np.empty(m.size)
    nest[ring2nest(nside_in, np.arange(m.size))] = m
    fac = (nside_in // nside_out) ** 2
    out_nest = nest.reshape(-1, fac).sum(axis=1)
    if power is None:
        out_nest /= fac
    elif power != -2:
        out_nest *= (float(nside_out) / nside_in) ** power / fac
    return out_nest[ring2nest(nside_out, np.arange(12 * nside_out ** 2))]


# ------------------------------------------------------------------ file I/O (.npy container)
def _is_fits_name(filename):
    """FITS for `.fits` names (what a stock plancklens / healpy expects to find); PLK_CACHE_FORMAT=npy keeps the
    `.npy` containers of the throughput runs (60 x faster to write for an lmax-2048 alm)."""
    import os
    if os.environ.get('PLK_CACHE_FORMAT', 'fits').lower() == 'npy':
        return False
    f = str(filename)
    return f.endswith('.fits') or f.endswith('.fits.gz') or f.endswith('.fit')


def _is_npy(filename):
    import gzip
    op = gzip.open if str(filename).endswith('.gz') else open
    try:
        with op(filename, 'rb') as f:
            return f.read(6) == b'\x93NUMPY'
    except OSError:
        with open(filename, 'rb') as f:          # '.gz' name holding a plain .npy (caches of earlier versions)
            return f.read(6) == b'\x93NUMPY'


def write_alm(filename, alms, out_dtype=None, lmax=-1, mmax=-1, mmax_in=-1, overwrite=True, **kw):
    """healpy.write_alm: FITS binary table (index, real, imag) for `.fits` names -- the files a stock plancklens /
    healpy reads (`fitsio.py`) -- a `.npy` container otherwise."""
    import os
    if not overwrite and os.path.exists(filename):
        raise OSError('File exists: %s' % filename)
    if _is_fits_name(filename):
        from . import fitsio
        return fitsio.write_alm(filename, alms, lmax=lmax, mmax=mmax, out_dtype=np.float64 if out_dtype is None else out_dtype)
    with open(filename, 'wb') as f:
        np.save(f, np.asarray(alms))


def read_alm(filename, hdu=1, return_mmax=False):
    if _is_npy(filename):
        with open(filename, 'rb') as f:
            a = np.load(f)
        return (a, Alm.getlmax(a.size)) if return_mmax else a
    from . import fitsio
    return fitsio.read_alm(filename, hdu=hdu, return_mmax=return_mmax)


def write_map(filename, m, nest=False, dtype=None, coord=None, overwrite=True, **kw):
    """healpy.write_map: FITS binary table (1024 pixels per row) for `.fits` / `.fits.gz` names, `.npy` otherwise."""
    import os
    if not overwrite and os.path.exists(filename):
        raise OSError('File exists: %s' % filename)
    if _is_fits_name(filename):
        from . import fitsio
        return fitsio.write_map(filename, m, nest=nest, dtype=np.float64 if dtype is None else dtype, coord=coord)
    with open(filename, 'wb') as f:
        np.save(f, np.asarray(m))


def read_map(filename, field=0, nest=False, hdu=1, **kw):
    """healpy.read_map: returns RING-ordered map(s) unless nest=True (or nest=None: as stored)."""
    if _is_npy(filename):
        with open(filename, 'rb') as f:
            m = np.load(f)
        if m.ndim == 2:
            return m[field] if np.isscalar(field) else m[list(field)]
        return m
    from . import fitsio
    m, hdr = fitsio.read_map(filename, field=field, hdu=hdu, return_header=True)
    stored_nest = str(hdr.get('ORDERING', 'RING')).strip().upper().startswith('NEST')
    if nest is None or stored_nest == bool(nest):
        return m
    maps = [m] if np.isscalar(field) else m
    nside = npix2nside(maps[0].size)
    r2n = ring2nest(nside, np.arange(maps[0].size))
    if stored_nest:                       # NEST on disk -> RING
        maps = [x[r2n] for x in maps]
    else:                                 # RING on disk -> NEST
        out = []
        for x in maps:
            y = np.empty_like(x)
            y[r2n] = x
            out.append(y)
        maps = out
    return maps[0] if np.isscalar(field) else maps


# ------------------------------------------------------------------ transforms (GPU)
def _plan(nside, lmax):
    from . import sht
    return sht.get_plan(nside, lmax)


def alm2map(alms, nside, lmax=None, mmax=None, pol=True, **kw):
    """Scalar synthesis (healpy signature).  reference: plancklens/shts.py:35."""
    if isinstance(alms, (list, tuple)) or np.ndim(alms) == 2:
        # (tlm, elm, blm) with pol=True -> [T, Q, U]: a spin-0 and a spin-2 synthesis in the HEALPix polarization
        # convention (reference use: qcinv/opfilt_tp.py:279)
        assert len(alms) == 3 and pol, "a list of alms is a (tlm, elm, blm) triple with pol=True"
        L = Alm.getlmax(np.asarray(alms[0]).size) if lmax is None else lmax
        q, u = alm2map_spin([alms[1], alms[2]], nside, 2, L)
        return [alm2map(alms[0], nside, lmax=L), q, u]
    alms = np.asarray(alms)
    assert alms.ndim == 1
    if lmax is None:
        lmax = Alm.getlmax(alms.size)
    assert mmax is None or mmax == lmax
    if alms.size != Alm.getsize(lmax):
        raise TypeError("Wrong alm size for the given lmax")
    return _plan(nside, lmax).alm2map_host(0, alms)


def map2alm(maps, lmax=None, mmax=None, iter=3, pol=True, use_weights=False, **kw):
    """Scalar analysis with uniform pixel weights 4 pi / npix and `iter` Jacobi refinement passes
    alm <- alm + map2alm(map - alm2map(alm)), healpy's signature and default (iter=3).  Every call on the reference's
    hot path passes iter=0 (plancklens/shts.py:35, opfilt_tt.py:34).  A (T, Q, U) triple with pol=True gives
    [tlm, elm, blm].  Ring-weight files are not shipped: use_weights=True raises."""
    if use_weights:
        raise NotImplementedError("map2alm(use_weights=True) needs healpy's ring-weight data files")
    if isinstance(maps, (list, tuple)) or np.ndim(maps) == 2:
        assert len(maps) == 3 and pol, "a list of maps is a (T, Q, U) triple with pol=True"
        tlm = map2alm(maps[0], lmax=lmax, mmax=mmax, iter=iter)
        elm, blm = _map2alm_spin_iter(maps[1], maps[2], 2, lmax, iter)
        return [tlm, elm, blm]
    m = np.asarray(maps, dtype=float)
    assert m.ndim == 1
    nside = npix2nside(m.size)
    if lmax is None:
        lmax = 3 * nside - 1
    assert mmax is None or mmax == lmax
    plan = _plan(nside, lmax)
    if iter == 0:
        return plan.map2alm_host(0, m)
    from . import sht
    md = sht.dev_map(m)
    alm = plan.map2alm(md)
    for _ in range(iter):
        res = plan.alm2map(alm)
        res = md - res
        sht.alm_axpy(alm, plan.map2alm(res), 1.0)
    return alm.cpu().numpy()


def _map2alm_spin_iter(m1, m2, spin, lmax, iter):
    m1, m2 = np.asarray(m1, dtype=float), np.asarray(m2, dtype=float)
    nside = npix2nside(m1.size)
    if lmax is None:
        lmax = 3 * nside - 1
    plan = _plan(nside, lmax)
    if iter == 0:
        return plan.map2alm_host(spin, m1, m2)
    from . import sht
    d1, d2 = sht.dev_map(m1), sht.dev_map(m2)
    g, c = plan.map2alm_spin(d1, d2, spin)
    for _ in range(iter):
        r1, r2 = plan.alm2map_spin(g, c, spin)
        dg, dc = plan.map2alm_spin(d1 - r1, d2 - r2, spin)
        sht.alm_axpy(g, dg, 1.0)
        sht.alm_axpy(c, dc, 1.0)
    return g.cpu().numpy(), c.cpu().numpy()


def smoothing(map_in, fwhm=0.0, sigma=None, beam_window=None, pol=False, iter=3, lmax=None, mmax=None,
              use_weights=False, **kw):
    """Gaussian (or `beam_window`) smoothing of a scalar map in harmonic space: map2alm (iter refinement passes, healpy's
    default 3) -> almxfl -> alm2map.  reference use: utils.apodize_mask (utils.py:296, :301)."""
    m = np.asarray(map_in, dtype=float)
    assert m.ndim == 1 and not pol, 'scalar maps only'
    nside = npix2nside(m.size)
    if lmax is None:
        lmax = 3 * nside - 1
    if sigma is None:
        sigma = fwhm / np.sqrt(8.0 * np.log(2.0))
    if beam_window is None:
        l = np.arange(lmax + 1, dtype=float)
        beam_window = np.exp(-0.5 * l * (l + 1) * sigma ** 2)
    alm = map2alm(m, lmax=lmax, mmax=mmax, iter=iter, use_weights=use_weights)
    return alm2map(almxfl(alm, beam_window), nside, lmax=lmax)


def alm2map_spin(alms, nside, spin, lmax, mmax=None):
    if spin <= 0 or spin > 3:
        raise ValueError("spin must be 1, 2 or 3")
    assert mmax is None or mmax == lmax
    a1 = np.asarray(alms[0])
    a2 = np.asarray(alms[1])
    if a1.size != Alm.getsize(lmax) or a2.size != a1.size:
        raise TypeError("Wrong alm size for the given lmax")
    m1, m2 = _plan(nside, lmax).alm2map_host(spin, a1, a2)
    return [m1, m2]


def map2alm_spin(maps, spin, lmax=None, mmax=None):
    if spin <= 0 or spin > 3:
        raise ValueError("spin must be 1, 2 or 3")
    m1 = np.asarray(maps[0], dtype=float)
    m2 = np.asarray(maps[1], dtype=float)
    nside = npix2nside(m1.size)
    assert m2.size == m1.size
    if lmax is None:
        lmax = 3 * nside - 1
    assert mmax is None or mmax == lmax
    a1, a2 = _plan(nside, lmax).map2alm_host(spin, m1, m2)
    return [a1, a2]


def install_as_healpy(force=False):
    """Registers this module as `healpy` in `sys.modules` when no real healpy can be imported, so that a parameter file
    written for the reference (`import healpy as hp`, reference params/idealized_example.py:24) runs unchanged on a
    machine without healpy.  A real healpy is never shadowed unless `force`.  Returns the module now serving `healpy`."""
    import importlib.util
    import sys
    if not force:
        if 'healpy' in sys.modules:
            return sys.modules['healpy']
        try:
            if importlib.util.find_spec('healpy') is not None:
                import healpy
                return healpy
        except (ImportError, ValueError):
            pass
    sys.modules['healpy'] = sys.modules[__name__]
    return sys.modules['healpy']
