"""Host-side utilities with the reference's names (reference: plancklens/utils.py).  numpy only."""
import hashlib
import sys
import time

import numpy as np


def alm_copy(alm, lmax=None):
    """Copy of a triangular alm array, optionally truncated to a lower lmax (reference: utils.py:19)."""
    alm = np.asarray(alm)
    lmax_in = int(np.floor(np.sqrt(2 * alm.size) - 1))
    if lmax is None or lmax == lmax_in:
        return alm.copy()
    assert lmax <= lmax_in, (lmax, lmax_in)
    out = np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=complex)
    for m in range(lmax + 1):
        o = m * (2 * lmax + 1 - m) // 2
        i = m * (2 * lmax_in + 1 - m) // 2
        out[o + m:o + lmax + 1] = alm[i + m:i + lmax + 1]
    return out


def clhash(cl, dtype=np.float16):
    """sha1 of an array cast to low precision (reference: utils.py:115)."""
    return hashlib.sha1(np.copy(np.asarray(cl).astype(dtype), order='C')).hexdigest()


def mchash(cl):
    return hashlib.sha1(np.copy(np.sort(cl), order='C')).hexdigest()


def cli(cl):
    """Pseudo-inverse of a non-negative array (reference: utils.py:132)."""
    cl = np.asarray(cl)
    out = np.zeros_like(cl)
    pos = cl > 0
    out[pos] = 1.0 / cl[pos]
    return out


def joincls(cls_list):
    n = min(len(c) for c in cls_list)
    return np.prod(np.array([c[:n] for c in cls_list]), axis=0)


def extcl(lmax, cl):
    out = np.zeros(lmax + 1)
    n = min(len(cl), lmax + 1)
    out[:n] = cl[:n]
    return out


def enumerate_progress(items, label=''):
    t0 = time.time()
    n = len(items)
    for i, v in enumerate(items):
        yield i, v
        if n and int(100.0 * i / n) > int(100.0 * (i - 1) / n):
            dt = time.time() - t0
            sys.stdout.write("\r [%02d:%02d:%02d] %s %s> %02d%%" % (dt // 3600, (dt % 3600) // 60, dt % 60, label,
                                                                  "-" * int(10.0 * i / n), int(100.0 * i / n)))
            sys.stdout.flush()
    sys.stdout.write("\n")


def hash_check(hash1, hash2, ignore=('lib_dir', 'prefix'), keychain=(), fn=None):
    """Raises AssertionError when two (nested) hash dictionaries differ (reference: utils.py:144)."""
    k1 = [k for k in hash1.keys() if k not in ignore]
    k2 = [k for k in hash2.keys() if k not in ignore]
    for key in set(k1).union(k2):
        if key not in hash1 or key not in hash2:
            raise KeyError("Cannot find key %s in hashdict %s" % (key, fn))
        v1, v2 = hash1[key], hash2[key]
        where = "HASHCHECK FAIL AT KEY %s (%s) in %s" % (key, '/'.join(map(str, keychain)), fn)
        assert type(v1) == type(v2), where + ': unequal types'
        if isinstance(v1, dict):
            hash_check(v1, v2, ignore=ignore, keychain=tuple(keychain) + (key,), fn=fn)
        elif isinstance(v1, np.ndarray):
            assert np.allclose(v1, v2), where + ': unequal arrays'
        else:
            assert v1 == v2, where + ': %s vs %s' % (v1, v2)


def camb_clfile(fname, lmax=None):
    """CAMB lensedCls / lenspotentialCls text file -> dict of C_l arrays (reference: utils.py:308)."""
    cols = np.loadtxt(fname).transpose()
    ell = cols[0].astype(int)
    if lmax is None:
        lmax = ell[-1]
    assert ell[-1] >= lmax, (ell[-1], lmax)
    sel = ell <= lmax
    w = ell * (ell + 1) / (2.0 * np.pi)
    cls = {}
    for i, k in enumerate(['tt', 'ee', 'bb', 'te']):
        cls[k] = np.zeros(lmax + 1)
        cls[k][ell[sel]] = cols[i + 1][sel] / w[sel]
    if len(cols) > 5:
        e = ell[sel].astype(float)
        cls['pp'] = np.zeros(lmax + 1); cls['pt'] = np.zeros(lmax + 1); cls['pe'] = np.zeros(lmax + 1)
        cls['pp'][ell[sel]] = cols[5][sel] / (e ** 2 * (e + 1) ** 2 / (2.0 * np.pi))
        cls['pt'][ell[sel]] = cols[6][sel] / (np.sqrt(e ** 3 * (e + 1) ** 3) / (2.0 * np.pi))
        cls['pe'][ell[sel]] = cols[7][sel] / (np.sqrt(e ** 3 * (e + 1) ** 3) / (2.0 * np.pi))
    return cls


def synthetic_cls(lmax):
    """Smooth CMB-like fiducial spectra used when no CAMB file is at hand (tests, smoke).  Not physical."""
    l = np.arange(lmax + 1, dtype=float)
    lp = np.maximum(l, 2.0)
    tt = 6e3 * (2 * np.pi) / (lp * (lp + 1)) * np.exp(-(lp / 1400.0) ** 1.3) * (1 + 0.3 * np.cos(lp / 95.0))
    ee = 40.0 * (2 * np.pi) / (lp * (lp + 1)) * (lp / 1000.0) ** 2 / (1 + (lp / 900.0) ** 4) * (1 + 0.4 * np.sin(lp / 95.0)) + 1e-7 * tt
    bb = 0.08 * (2 * np.pi) / (lp * (lp + 1)) * (lp / 1000.0) ** 2 / (1 + (lp / 1500.0) ** 4)
    te = 0.5 * np.sqrt(tt * ee) * np.cos(lp / 60.0)
    for c in (tt, ee, bb, te):
        c[:2] = 0.0
    return {'tt': tt, 'ee': ee, 'bb': bb, 'te': te}


def cl_inverse(cls):
    """Inverse of the T, E, B spectral matrices; dictionaries in and out (reference: utils.py:336-366)."""
    def ext(cl, lmax):
        ret = np.zeros(lmax + 1, dtype=float)
        n = min(len(cl), lmax + 1)
        ret[:n] = np.asarray(cl)[:n]
        return ret
    lmax = np.max([len(cl) for cl in cls.values()]) - 1
    m = np.zeros((lmax + 1, 3, 3))
    for k, (i, j) in zip(['tt', 'ee', 'bb', 'te', 'tb', 'eb'], [[0, 0], [1, 1], [2, 2], [0, 1], [0, 2], [1, 2]]):
        m[:, i, j] = m[:, j, i] = ext(cls.get(k, [0.]), lmax)
    mi = np.linalg.pinv(m)
    out = {}
    for k, (i, j) in zip(['tt', 'ee', 'bb', 'te', 'tb', 'eb'], [[0, 0], [1, 1], [2, 2], [0, 1], [0, 2], [1, 2]]):
        arr = mi[:, i, j].copy()
        if np.any(arr):
            out[k] = arr
    return out


_TEB = ('t', 'e', 'b')


def _cldict2arr(cls_dict):
    """{'tt': ..., 'te': ...} -> symmetric (3, 3, lmax + 1) array in T, E, B order (reference: utils.py:375-381)."""
    n = max(len(cl) for cl in cls_dict.values())
    out = np.zeros((3, 3, n), dtype=float)
    for i, x in enumerate(_TEB):
        for j, y in enumerate(_TEB):
            cl = cls_dict.get(x + y, cls_dict.get(y + x, None))
            if cl is not None:
                out[i, j] = extcl(n - 1, np.asarray(cl, dtype=float))
    return out


def cls_dot(cls_list, ret_dict=False):
    """Product, per multipole, of T E B spectral matrices given as dictionaries or (3, 3, lmax + 1) arrays
    (reference: utils.py:383-410).  With `ret_dict` the non-zero entries of the upper triangle come back as a dict."""
    mats = [_cldict2arr(c) if isinstance(c, dict) else np.asarray(c) for c in cls_list]
    ret = mats[-1]
    for m in mats[-2::-1]:
        ret = np.einsum('ikl,kjl->ijl', m, ret)
    if not ret_dict or len(mats) == 1:
        return ret
    out = {}
    for k, (i, j) in zip(['tt', 'ee', 'bb', 'te', 'tb', 'eb'], [[0, 0], [1, 1], [2, 2], [0, 1], [0, 2], [1, 2]]):
        if np.any(ret[i, j]):
            out[k] = ret[i, j].copy()
    return out


def alm2rlm(alm):
    """complex alm -> 'real harmonic' coefficients (reference: utils.py:37-51)"""
    from .qcinv.dense import alm2rlm as f
    return f(alm)


def rlm2alm(rlm):
    """inverse of alm2rlm (reference: utils.py:54-69)"""
    from .qcinv.dense import rlm2alm as f
    return f(rlm)


class stats:
    """Running mean and covariance of data vectors (reference: utils.py:178-267).

    `sum`, `mom` (second-moment matrix, only with docov) and `N` are the accumulators; unlike the reference the diagonal
    second moments are always kept, so `sigmas()` also works with docov=False."""

    def __init__(self, size, xcoord=None, docov=True):
        self.N = 0
        self.size = size
        self.sum = np.zeros(size)
        self.sum2 = np.zeros(size)
        if docov:
            self.mom = np.zeros((size, size))
        self.xcoord = xcoord
        self.docov = docov

    def add(self, v):
        assert v.shape == (self.size,), "input not understood"
        self.sum += v
        self.sum2 += v * v
        if self.docov:
            self.mom += np.outer(v, v)
        self.N += 1

    def mean(self):
        assert self.N > 0
        return self.sum / float(self.N)

    avg = mean

    def cov(self):
        assert self.docov and self.N > 0
        if self.N == 1:
            return np.zeros((self.size, self.size))
        m = self.mean()
        return (self.mom - self.N * np.outer(m, m)) / (self.N - 1.)

    def sigmas(self):
        if self.docov:
            return np.sqrt(np.diagonal(self.cov()))
        assert self.N > 0
        if self.N == 1:
            return np.zeros(self.size)
        return np.sqrt(np.maximum(self.sum2 - self.N * self.mean() ** 2, 0.) / (self.N - 1.))

    def corrcoeffs(self):
        sig = self.sigmas()
        return self.cov() / np.outer(sig, sig)

    def sigmas_on_mean(self):
        assert self.N > 0
        return self.sigmas() / np.sqrt(self.N)

    def inverse(self, bias_p=None):
        """inverse covariance, by default with the (N - size - 2) / (N - 1) de-biasing factor"""
        assert self.N > self.size, "Non invertible cov.matrix"
        if bias_p is None:
            bias_p = (self.N - self.size - 2.) / (self.N - 1)
        return bias_p * np.linalg.inv(self.cov())

    def get_chisq(self, data):
        assert data.size == self.size, (data.size, self.size)
        dx = data - self.mean()
        return float(dx @ self.inverse() @ dx)

    def get_chisq_pte(self, data):
        from scipy.stats import chi2
        return chi2.sf(self.get_chisq(data), self.N - 1)

    def rebin_that_nooverlap(self, orig_coord, lmins, lmaxs, weights=None):
        """New instance with the data vector averaged (weights) inside the non-overlapping bins [lmin, lmax]."""
        lmins, lmaxs = np.asarray(lmins), np.asarray(lmaxs)
        assert orig_coord.size == self.size and lmins.size == lmaxs.size, "Incompatible input"
        assert np.all(np.diff(lmins) > 0.) and np.all(np.diff(lmaxs) > 0.), "This only for non overlapping bins."
        if weights is None:
            weights = np.ones(self.size)
        assert weights.size == self.size and self.size > len(lmaxs), "incompatible input"
        T = np.zeros((len(lmaxs), self.size))
        for k, (lo, hi) in enumerate(zip(lmins, lmaxs)):
            sel = (orig_coord >= lo) & (orig_coord <= hi)
            if np.any(sel):
                T[k, sel] = weights[sel] / np.sum(weights[sel])
        new = stats(len(lmaxs), xcoord=0.5 * (lmins[:-1] + lmaxs[1:]), docov=self.docov)
        new.sum = T @ self.sum
        if self.docov:
            new.mom = T @ self.mom @ T.T
            new.sum2 = np.diagonal(new.mom).copy()
        new.N = self.N
        return new


def apodize_mask(mask, sigma_arcmin=12., lmax=None, method='hybrid', cache_dir='caches/', mult_factor=3, min_factor=0.1):
    """Apodizes a mask for pseudo-C_l work (reference: utils.py:270-306): Gaussian smoothing, or 'hybrid' -- smooth, scale
    (1 - mask) by `mult_factor`, clip, smooth again with half the width, which mostly eats into the unmasked side.
    The smoothing transforms run on the GPU (`hp.smoothing`)."""
    import hashlib
    import os
    from . import hp
    if not sigma_arcmin:
        return mask
    sigma = sigma_arcmin / 180. / 60. * np.pi
    name = None
    if cache_dir:
        tag = '_'.join('%s' % x for x in [sigma_arcmin, method, lmax, mult_factor, min_factor, hashlib.sha1(mask).hexdigest()])
        name = os.path.join(cache_dir, 'ap_mask_' + tag) + '.fits'
        if os.path.exists(name):
            return hp.read_map(name)
    print('apodizing... (fsky_unapodized=%s)' % (np.sum(mask ** 2) / mask.size))
    ap = hp.smoothing(mask, sigma=sigma, lmax=lmax)
    if method == 'gaussian':
        return ap
    if method != 'hybrid':
        raise ValueError('Unknown apodization method')
    ap = 1 - np.minimum(1., np.maximum(0., mult_factor * (1 - ap) - min_factor))
    ap = hp.smoothing(ap, sigma=sigma / 2, lmax=lmax)
    print('fsky=', np.sum(ap ** 2) / ap.size)
    if name is not None:
        os.makedirs(cache_dir, exist_ok=True)
        hp.write_map(name, ap)
    return ap
