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
