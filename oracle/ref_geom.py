"""HEALPix RING geometry and alm index helpers (oracle side, test infrastructure only).

Restates the conventions the reference relies on through healpy
(`hp.nside2npix`, `hp.Alm.getidx/getsize/getlmax`, ring layout used by `hp.alm2map`),
see SURVEY.md section 8(c) "Definitions the restatement must implement".
"""
import numpy as np


def nside2npix(nside):
    return 12 * nside * nside


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    assert 12 * nside * nside == npix, npix
    return nside


def alm_getsize(lmax, mmax=None):
    mmax = lmax if mmax is None else mmax
    return mmax * (2 * lmax + 1 - mmax) // 2 + lmax + 1


def alm_getlmax(size):
    lmax = int(np.floor(np.sqrt(2 * size) - 1))
    if (lmax + 1) * (lmax + 2) // 2 != size:
        return -1
    return lmax


def alm_getidx(lmax, l, m):
    return m * (2 * lmax + 1 - m) // 2 + l


def ring_info(nside):
    """Per-ring (1..4N-1) arrays: nphi, startpix, z=cos(theta), sin(theta), phi0."""
    N = nside
    i = np.arange(1, 4 * N)
    nphi = np.empty(i.size, dtype=np.int64)
    start = np.empty(i.size, dtype=np.int64)
    z = np.empty(i.size)
    sth = np.empty(i.size)
    phi0 = np.empty(i.size)
    npix = 12 * N * N
    for k, ir in enumerate(i):
        if ir < N:
            u = ir * ir / (3.0 * N * N)
            nphi[k] = 4 * ir
            start[k] = 2 * ir * (ir - 1)
            z[k] = 1.0 - u
            sth[k] = np.sqrt(u * (2.0 - u))
            phi0[k] = np.pi / (4.0 * ir)
        elif ir <= 3 * N:
            nphi[k] = 4 * N
            start[k] = 2 * N * (N - 1) + (ir - N) * 4 * N
            z[k] = (2 * N - ir) * 2.0 / (3.0 * N)
            sth[k] = np.sqrt((1.0 - z[k]) * (1.0 + z[k]))
            phi0[k] = np.pi / (4.0 * N) if ((ir - N) % 2 == 0) else 0.0
        else:
            ip = 4 * N - ir
            u = ip * ip / (3.0 * N * N)
            nphi[k] = 4 * ip
            start[k] = npix - 2 * ip * (ip + 1)
            z[k] = -(1.0 - u)
            sth[k] = np.sqrt(u * (2.0 - u))
            phi0[k] = np.pi / (4.0 * ip)
    return nphi, start, z, sth, phi0


def pix2ang(nside):
    """theta, phi of every RING-ordered pixel."""
    nphi, start, z, sth, phi0 = ring_info(nside)
    theta = np.empty(12 * nside * nside)
    phi = np.empty(12 * nside * nside)
    for k in range(nphi.size):
        sl = slice(start[k], start[k] + nphi[k])
        theta[sl] = np.arctan2(sth[k], z[k])
        phi[sl] = phi0[k] + 2.0 * np.pi * np.arange(nphi[k]) / nphi[k]
    return theta, phi


# ---------------------------------------------------------------- RING <-> NEST
def _spread_bits(v):
    v = v.astype(np.int64)
    v = (v | (v << 16)) & 0x0000FFFF0000FFFF
    v = (v | (v << 8)) & 0x00FF00FF00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v << 2)) & 0x3333333333333333
    v = (v | (v << 1)) & 0x5555555555555555
    return v


_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def ring2nest(nside, ipring):
    """Standard HEALPix ring->nest index map (nside power of two)."""
    N = nside
    ip = np.asarray(ipring, dtype=np.int64)
    npix = 12 * N * N
    ncap = 2 * N * (N - 1)
    iring = np.empty_like(ip)
    iphi = np.empty_like(ip)
    kshift = np.empty_like(ip)
    nr = np.empty_like(ip)
    face = np.empty_like(ip)
    north = ip < ncap
    south = ip >= npix - ncap
    equat = ~(north | south)
    # north cap
    p = ip[north]
    ir = ((1 + np.floor(np.sqrt(1 + 2 * p.astype(np.float64))).astype(np.int64)) >> 1)
    ir = np.where(2 * ir * (ir - 1) > p, ir - 1, ir)
    ir = np.where(2 * ir * (ir + 1) <= p, ir + 1, ir)
    iring[north] = ir
    iphi[north] = p + 1 - 2 * ir * (ir - 1)
    kshift[north] = 0
    nr[north] = ir
    face[north] = (iphi[north] - 1) // ir
    # equatorial
    p = ip[equat] - ncap
    ir = p // (4 * N) + N
    iring[equat] = ir
    iphi[equat] = p % (4 * N) + 1
    kshift[equat] = (ir + N) & 1
    nr[equat] = N
    ire = ir - N + 1
    irm = 2 * N + 2 - ire
    ifm = (iphi[equat] - ire // 2 + N - 1) // N
    ifp = (iphi[equat] - irm // 2 + N - 1) // N
    face[equat] = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))
    # south cap
    p = npix - ip[south]
    ir = ((1 + np.floor(np.sqrt(2 * p.astype(np.float64) - 1)).astype(np.int64)) >> 1)
    ir = np.where(2 * ir * (ir - 1) >= p, ir - 1, ir)
    ir = np.where(2 * ir * (ir + 1) < p, ir + 1, ir)
    iphi[south] = 4 * ir + 1 - (p - 2 * ir * (ir - 1))
    kshift[south] = 0
    nr[south] = ir
    face[south] = 8 + (iphi[south] - 1) // ir
    iring[south] = 4 * N - ir
    irt = iring - _JRLL[face] * N + 1
    ipt = 2 * iphi - _JPLL[face] * nr - kshift - 1
    ipt = np.where(ipt >= 2 * N, ipt - 8 * N, ipt)
    ix = (ipt - irt) >> 1
    iy = (-(ipt + irt)) >> 1
    return face * N * N + _spread_bits(ix) + (_spread_bits(iy) << 1)


def ud_grade_sum(m, nside_out):
    """healpy.ud_grade(m, nside_out, power=-2) for nside_out < nside_in, RING in / RING out:
    every output pixel is the SUM of its (nside_in/nside_out)^2 children
    (reference call sites: opfilt_tt.py:179, opfilt_pp.py:251)."""
    nside_in = npix2nside(len(m))
    assert nside_out <= nside_in and nside_in % nside_out == 0
    if nside_out == nside_in:
        return np.array(m, dtype=float, copy=True)
    r2n_in = ring2nest(nside_in, np.arange(len(m)))
    nest_in = np.empty(len(m))
    nest_in[r2n_in] = m
    fac = (nside_in // nside_out) ** 2
    nest_out = nest_in.reshape(-1, fac).sum(axis=1)
    r2n_out = ring2nest(nside_out, np.arange(12 * nside_out ** 2))
    return nest_out[r2n_out]
