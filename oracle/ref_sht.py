"""CPU oracle SHT with healpy-compatible signatures (test infrastructure only; parity unpinned, see
oracle/__init__.py).  Legendre stage: oracle/csht.c through ctypes.  Ring FFT stage: numpy (pocketfft).

Replaces, for tests and the CPU baseline only, the healpy calls of the reference hot path:
`hp.alm2map`, `hp.map2alm(iter=0)`, `hp.alm2map_spin`, `hp.map2alm_spin`
(/root/reference/plancklens/shts.py:33-35, utils_spin.py:21-34).
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.fft as sfft

from . import ref_geom as rg

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/csht.c -> oracle/_build/libcsht.so (gcc + OpenMP)."""
    so = os.path.join(_HERE, '_build', 'libcsht.so')
    src = os.path.join(_HERE, 'csht.c')
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, '_build', 'libcsht.so')
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        vp = ctypes.c_void_p
        ldp = ctypes.POINTER(ctypes.c_longdouble)
        L.csht_synth.argtypes = [ctypes.c_int] * 4 + [ip, ip, dp, ldp, ldp, vp, vp, vp, vp, ctypes.c_int]
        L.csht_anal.argtypes = [ctypes.c_int] * 4 + [ip, ip, dp, ldp, ldp, dp, vp, vp, vp, vp, ctypes.c_int]
        L.csht_synth_mlist.argtypes = [ctypes.c_int] * 4 + [ip, ip, dp, ldp, ldp, vp, vp, vp, vp, ip, ctypes.c_int]
        L.csht_anal_mlist.argtypes = [ctypes.c_int] * 4 + [ip, ip, dp, ldp, ldp, dp, vp, vp, vp, vp, ip, ctypes.c_int]
        L.csht_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def max_threads():
    return _lib().csht_max_threads()


class _Geom:
    _cache = {}

    def __init__(self, nside):
        N = nside
        self.nside = N
        self.nphi, self.start, self.z, self.sth, self.phi0 = rg.ring_info(N)
        self.nring = 4 * N - 1
        self.npair = 2 * N
        i = np.arange(1, 2 * N + 1)
        self.rn = (i - 1).astype(np.int32)
        self.rs = np.where(i < 2 * N, 4 * N - i - 1, -1).astype(np.int32)
        # ring-pair geometry in x87 long double from the exact integer formulas: 1-z = i^2/(3N^2) in the cap,
        # z = 2(2N-i)/(3N) in the belt.  log sin/cos(theta/2) are handed to C in long double so that the
        # sin^m seed carries no m*eps amplification of a double-rounded angle.
        il = i.astype(np.longdouble)
        NL = np.longdouble(N)
        omz = np.where(i < N, il * il / (3 * NL * NL), 1 - 2 * (2 * NL - il) / (3 * NL))
        self.cth = np.ascontiguousarray((1 - omz).astype(np.float64))
        self.lshalf = np.ascontiguousarray(0.5 * np.log(omz / 2))
        self.lchalf = np.ascontiguousarray(0.5 * np.log(1 - omz / 2))
        self.weight = np.full(self.npair, 4.0 * np.pi / (12 * N * N))

    @classmethod
    def get(cls, nside):
        if nside not in cls._cache:
            cls._cache[nside] = cls(nside)
        return cls._cache[nside]


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def legendre_synth(nside, spin, lmax, mmax, almG, almC=None, mstep=1):
    """-> phase arrays X1 (, X2) of shape [nring, mmax+1]."""
    g = _Geom.get(nside)
    almG = np.ascontiguousarray(almG, dtype=np.complex128)
    X1 = np.zeros((g.nring, mmax + 1), dtype=np.complex128)
    X2 = np.zeros((g.nring, mmax + 1), dtype=np.complex128) if spin > 0 else None
    if spin > 0:
        almC = np.ascontiguousarray(almC, dtype=np.complex128)
    _lib().csht_synth(spin, lmax, mmax, g.npair, _p(g.rn, ctypes.c_int), _p(g.rs, ctypes.c_int),
                      _p(g.cth, ctypes.c_double), _p(g.lchalf, ctypes.c_longdouble), _p(g.lshalf, ctypes.c_longdouble),
                      almG.ctypes.data, almC.ctypes.data if spin > 0 else None,
                      X1.ctypes.data, X2.ctypes.data if spin > 0 else None, mstep)
    return X1, X2


def legendre_anal(nside, spin, lmax, mmax, X1, X2=None, mstep=1):
    g = _Geom.get(nside)
    nalm = rg.alm_getsize(lmax, mmax)
    G = np.zeros(nalm, dtype=np.complex128)
    C = np.zeros(nalm, dtype=np.complex128) if spin > 0 else None
    X1 = np.ascontiguousarray(X1)
    if spin > 0:
        X2 = np.ascontiguousarray(X2)
    _lib().csht_anal(spin, lmax, mmax, g.npair, _p(g.rn, ctypes.c_int), _p(g.rs, ctypes.c_int),
                     _p(g.cth, ctypes.c_double), _p(g.lchalf, ctypes.c_longdouble), _p(g.lshalf, ctypes.c_longdouble),
                     _p(g.weight, ctypes.c_double), X1.ctypes.data, X2.ctypes.data if spin > 0 else None,
                     G.ctypes.data, C.ctypes.data if spin > 0 else None, mstep)
    return G, C


def legendre_synth_mlist(nside, spin, lmax, almG, almC, mlist):
    """Legendre synthesis of the m in `mlist` only -> COMPACT phase arrays X1 (, X2) of shape [nring, len(mlist)]
    (full-size parity samples: a handful of m, including m ~ lmax, cost seconds where the whole band costs minutes)."""
    g = _Geom.get(nside)
    ml = np.ascontiguousarray(mlist, dtype=np.int32)
    almG = np.ascontiguousarray(almG, dtype=np.complex128)
    X1 = np.zeros((g.nring, ml.size), dtype=np.complex128)
    X2 = np.zeros((g.nring, ml.size), dtype=np.complex128) if spin > 0 else None
    if spin > 0:
        almC = np.ascontiguousarray(almC, dtype=np.complex128)
    _lib().csht_synth_mlist(spin, lmax, lmax, g.npair, _p(g.rn, ctypes.c_int), _p(g.rs, ctypes.c_int),
                            _p(g.cth, ctypes.c_double), _p(g.lchalf, ctypes.c_longdouble), _p(g.lshalf, ctypes.c_longdouble),
                            almG.ctypes.data, almC.ctypes.data if spin > 0 else None,
                            X1.ctypes.data, X2.ctypes.data if spin > 0 else None, _p(ml, ctypes.c_int), ml.size)
    return X1, X2


def legendre_anal_mlist(nside, spin, lmax, X1, X2, mlist):
    """Legendre analysis of the m in `mlist` only, from COMPACT phase arrays [nring, len(mlist)] (weights 4 pi / npix
    applied here) -> full-size alm arrays, zero at every other m."""
    g = _Geom.get(nside)
    ml = np.ascontiguousarray(mlist, dtype=np.int32)
    nalm = rg.alm_getsize(lmax, lmax)
    G = np.zeros(nalm, dtype=np.complex128)
    C = np.zeros(nalm, dtype=np.complex128) if spin > 0 else None
    X1 = np.ascontiguousarray(X1, dtype=np.complex128)
    assert X1.shape == (g.nring, ml.size)
    if spin > 0:
        X2 = np.ascontiguousarray(X2, dtype=np.complex128)
    _lib().csht_anal_mlist(spin, lmax, lmax, g.npair, _p(g.rn, ctypes.c_int), _p(g.rs, ctypes.c_int),
                           _p(g.cth, ctypes.c_double), _p(g.lchalf, ctypes.c_longdouble), _p(g.lshalf, ctypes.c_longdouble),
                           _p(g.weight, ctypes.c_double), X1.ctypes.data, X2.ctypes.data if spin > 0 else None,
                           G.ctypes.data, C.ctypes.data if spin > 0 else None, _p(ml, ctypes.c_int), ml.size)
    return G, C


def phase2map(nside, X, workers=None):
    """Ring FFT stage of synthesis: X[ring, m] -> RING-ordered real map.
    map_j = X_0 + 2 Re sum_{m>0} X_m e^{i m (phi0 + 2 pi j / nphi)}; m >= nphi/2 aliases onto the ring's band.
    `workers`: threads of the batched pocketfft calls (scipy.fft; -1 = all cores), used by the CPU baseline."""
    g = _Geom.get(nside)
    mmax = X.shape[1] - 1
    out = np.empty(12 * nside * nside)
    m = np.arange(mmax + 1)

    def one(n):
        rows = np.where(g.nphi == n)[0]
        Xs = X[rows] * np.exp(1j * m[None, :] * g.phi0[rows][:, None])
        big = rows.size > 64                                      # the equatorial class: thread inside the batched FFT
        if 2 * mmax < n:
            # no aliasing: Hermitian half spectrum straight into a complex-to-real FFT
            h = np.zeros((rows.size, n // 2 + 1), dtype=complex)
            h[:, :mmax + 1] = Xs
            x = sfft.irfft(h, n=n, axis=1, workers=workers if big else None) * n
        else:
            full = np.zeros((rows.size, n), dtype=complex)       # spectrum of the m >= 0 part, wrapped mod n
            for j0 in range(0, mmax + 1, n):
                blk = Xs[:, j0:j0 + n]
                full[:, :blk.shape[1]] += blk
            d = full.copy()                                        # add the conjugate (m < 0) images
            d[:, 0] += np.conj(full[:, 0]) - np.conj(Xs[:, 0])    # m = 0 itself has no mirror term
            d[:, 1:] += np.conj(full[:, :0:-1])
            x = (sfft.ifft(d, axis=1, workers=workers if big else None) * n).real
        for a, r in enumerate(rows):
            out[g.start[r]:g.start[r] + n] = x[a]
    _for_each_ring_length(one, np.unique(g.nphi)[::-1], workers)
    return out


def _for_each_ring_length(fun, lengths, workers):
    """the ~nside distinct cap-ring lengths are independent: spread them over threads when the caller asks for any"""
    if workers in (None, 0, 1):
        for n in lengths:
            fun(n)
        return
    import concurrent.futures as cf
    import os
    nthr = os.cpu_count() if workers < 0 else workers
    with cf.ThreadPoolExecutor(max_workers=nthr) as pool:
        list(pool.map(fun, lengths))


def map2phase(nside, mp, mmax, workers=None):
    """Adjoint ring FFT: X[ring, m] = sum_j map_j e^{-i m phi_j} (no weights)."""
    g = _Geom.get(nside)
    X = np.empty((g.nring, mmax + 1), dtype=complex)
    m = np.arange(mmax + 1)

    def one(n):
        rows = np.where(g.nphi == n)[0]
        x = np.stack([mp[g.start[r]:g.start[r] + n] for r in rows])
        h = sfft.rfft(x, axis=1, workers=workers if rows.size > 64 else None)
        if mmax <= n // 2:
            d = h[:, :mmax + 1]
        else:
            k = m % n
            d = np.where((k <= n // 2)[None, :], h[:, np.minimum(k, n // 2)], np.conj(h[:, np.minimum(n - k, n // 2)]))
        X[rows] = d * np.exp(-1j * m[None, :] * g.phi0[rows][:, None])
    _for_each_ring_length(one, np.unique(g.nphi)[::-1], workers)
    return X


# ------------------------------------------------------------------ healpy-compatible entry points
def alm2map(alm, nside, lmax=None, mmax=None, **kw):
    alm = np.asarray(alm)
    if lmax is None:
        lmax = rg.alm_getlmax(alm.size)
    mmax = lmax if mmax is None else mmax
    assert alm.size == rg.alm_getsize(lmax, mmax), (alm.size, lmax, mmax)
    X1, _ = legendre_synth(nside, 0, lmax, mmax, alm)
    return phase2map(nside, X1)


def map2alm(m, lmax=None, mmax=None, iter=0, **kw):
    """Uniform-weight analysis; the reference hot path always passes iter=0.  iter > 0 restates the published
    refinement loop of HEALPix C++ `map2alm_iter` (alm_healpix_tools.cc; what healpy.map2alm(iter=k) runs):
    alm <- alm + map2alm(map - alm2map(alm)), k times.  Parity of this branch is unpinned (healpy absent)."""
    m = np.asarray(m, dtype=float)
    nside = rg.npix2nside(m.size)
    if lmax is None:
        lmax = 3 * nside - 1
    mmax = lmax if mmax is None else mmax

    def once(x):
        G, _ = legendre_anal(nside, 0, lmax, mmax, map2phase(nside, x, mmax))
        return G
    alm = once(m)
    for _ in range(iter):
        alm = alm + once(m - alm2map(alm, nside, lmax=lmax, mmax=mmax))
    return alm


def alm2map_spin(alms, nside, spin, lmax, mmax=None):
    assert spin > 0
    mmax = lmax if mmax is None else mmax
    X1, X2 = legendre_synth(nside, spin, lmax, mmax, alms[0], alms[1])
    return [phase2map(nside, X1), phase2map(nside, X2)]


def map2alm_spin(maps, spin, lmax=None, mmax=None):
    assert spin > 0
    m1 = np.asarray(maps[0], dtype=float)
    m2 = np.asarray(maps[1], dtype=float)
    nside = rg.npix2nside(m1.size)
    if lmax is None:
        lmax = 3 * nside - 1
    mmax = lmax if mmax is None else mmax
    X1 = map2phase(nside, m1, mmax)
    X2 = map2phase(nside, m2, mmax)
    G, C = legendre_anal(nside, spin, lmax, mmax, X1, X2)
    return [G, C]
