"""Preconditioned conjugate-directions solver (reference: plancklens/qcinv/cd_solve.py:35-107).

Same loop, same operation order (residual refresh every `roundoff` iterations, one cached previous direction
for `tr_cg`), written against vectors that live on the GPU: anything exposing `+ - *scalar` and, when
available, an in-place `axpy(a, x)`.
"""
import numpy as np


def PTR(p, t, r):
    return lambda i: max(0, i - max(p, int(min(t, np.mod(i, r)))))


tr_cg = (lambda i: i - 1)
tr_cd = (lambda i: 0)


class cache_mem(dict):
    """In-memory store of (dTAd_inv, searchdirs, searchfwds) per iteration."""

    def __init__(self):
        super().__init__()

    def store(self, key, data):
        self[key] = list(data)

    def restore(self, key):
        return self[key]

    def remove(self, key):
        del self[key]

    def trim(self, keys):
        keys = set(keys)
        assert keys.issubset(self.keys())
        for key in set(self.keys()) - keys:
            del self[key]


def _axpy(y, a, x):
    """y += a x, in place when the vector type supports it; returns the updated vector."""
    if hasattr(y, 'axpy'):
        return y.axpy(a, x)
    y += x * a
    return y


def cd_solve(x, b, fwd_op, pre_ops, dot_op, criterion, tr, cache=None, roundoff=25):
    """Solves x = fwd_op^{-1} b in place; returns the iteration count.

    fwd_op, pre_ops and dot_op must not modify their arguments (reference: cd_solve.py:51)."""
    cache = cache_mem() if cache is None else cache
    n_pre = len(pre_ops)

    residual = b - fwd_op(x)
    searchdirs = [op(residual) for op in pre_ops]

    it = 0
    while not criterion(it, x, residual):
        searchfwds = [fwd_op(sd) for sd in searchdirs]
        deltas = [dot_op(sd, residual) for sd in searchdirs]

        dTAd = np.zeros((n_pre, n_pre))
        for i1 in range(n_pre):
            for i2 in range(i1 + 1):
                dTAd[i1, i2] = dTAd[i2, i1] = dot_op(searchdirs[i1], searchfwds[i2])
        dTAd_inv = np.linalg.inv(dTAd)

        alphas = np.dot(dTAd_inv, deltas)
        for sd, alpha in zip(searchdirs, alphas):
            x = _axpy(x, alpha, sd)

        cache.store(it, [dTAd_inv, searchdirs, searchfwds])

        it += 1
        if np.mod(it, roundoff) == 0:
            residual = b - fwd_op(x)
        else:
            for sf, alpha in zip(searchfwds, alphas):
                residual = _axpy(residual, -alpha, sf)

        searchdirs = [pre_op(residual) for pre_op in pre_ops]

        # orthogonalise against the cached previous searches
        for titer in range(tr(it), it):
            prev_dTAd_inv, prev_dirs, prev_fwds = cache.restore(titer)
            for isd in range(n_pre):
                proj = [dot_op(searchdirs[isd], pf) for pf in prev_fwds]
                betas = np.dot(prev_dTAd_inv, proj)
                for beta, pd in zip(betas, prev_dirs):
                    searchdirs[isd] = _axpy(searchdirs[isd], -beta, pd)

        cache.trim(range(tr(it + 1), it))
    return it
