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


def can_solve_fixed(pre_ops, dot_op, tr, iter_max, eps_min):
    """True when cd_solve_fixed reproduces cd_solve: one preconditioner, standard PCG recurrence (tr_cg), a
    convergence test that can never fire (eps_min = 0) and a finite iteration count."""
    return (len(pre_ops) == 1 and tr is tr_cg and eps_min == 0.0 and np.isfinite(iter_max)
            and hasattr(dot_op, 'dev'))


def _comps(v):
    """device tensors of a dalm / eblm / teblm vector"""
    if hasattr(v, 't'):
        return [v]
    return [v.tlm, v.elm, v.blm] if hasattr(v, 'tlm') else [v.elm, v.blm]


def _update_pair(x, d, residual, Ad, alpha_dev):
    """x += alpha d ; residual -= alpha Ad: one kernel per component (plk_alm_axpy2_dev)"""
    from .. import sht
    for cx, cd, cr, ca in zip(_comps(x), _comps(d), _comps(residual), _comps(Ad)):
        sht.alm_axpy2(cx.t, cd.t, cr.t, ca.t, alpha_dev)
        cx.zero = cx.zero and cd.zero
        cr.zero = cr.zero and ca.zero


def cd_solve_fixed(x, b, fwd_op, pre_ops, dot_op, niter, roundoff=25):
    """cd_solve for the inner multigrid stages (iter_max iterations, eps_min = 0, tr_cg; reference
    multigrid.py:185-215 with the chains of filt_cinv.py:113-116): the same updates in the same order, with the
    step lengths alpha = (d.r)/(d.Ad) and beta = (d'.Ad)/(d.Ad) kept in device memory, so the whole solve is a fixed
    sequence of kernel launches with no host synchronisation -- it can be captured in a CUDA graph.  The residual-norm
    evaluations of the monitor (only logged, never acted on when eps_min = 0) and the search direction computed after
    the last update (never used) are skipped.

    Launches per iteration besides the two operators: three one-kernel dot products whose last block also forms the
    step length (`dot_op.fused`, plk_alm_dot_fused_dev), one fused solution + residual update per component
    (plk_alm_axpy2_dev) and one axpy for the new search direction."""
    from .. import sht
    (pre_op,) = pre_ops
    niter = int(niter)
    fused = hasattr(dot_op, 'fused')
    # A 0 = 0: skip the operator on the zero start vector the multigrid stages use (opfilt_tt.py:68 does the same)
    residual = b.copy() if (hasattr(x, 'is_zero') and x.is_zero()) else b - fwd_op(x)
    d = pre_op(residual)
    for it in range(1, niter + 1):
        Ad = fwd_op(d)
        if fused:
            delta = dot_op.fused(d, residual)                              # [d.r, -, -]
            t3 = dot_op.fused(d, Ad, num=delta[0:1])                       # [d.Ad, alpha, -alpha]
            dTAd, alpha, malpha = t3[0:1], t3[1:2], t3[2:3]
        else:
            delta = dot_op.dev(d, residual)
            dTAd = dot_op.dev(d, Ad)
            alpha, malpha = sht.scalar_ratio(delta, dTAd), None
        if it == niter:
            x = _axpy(x, alpha, d)
            break
        if it % roundoff == 0:
            x = _axpy(x, alpha, d)
            residual = b - fwd_op(x)
        elif fused:
            _update_pair(x, d, residual, Ad, alpha)
        else:
            x = _axpy(x, alpha, d)
            residual = _axpy(residual, sht.scalar_ratio(delta, dTAd, -1.0), Ad)
        dn = pre_op(residual)
        if fused:
            beta = dot_op.fused(dn, Ad, den=dTAd, scale=-1.0)[1:2]
        else:
            beta = sht.scalar_ratio(dot_op.dev(dn, Ad), dTAd, -1.0)
        d = _axpy(dn, beta, d)
    return niter


def can_solve_dev(pre_ops, dot_op, tr):
    """True when cd_solve_dev reproduces cd_solve: one preconditioner, the standard PCG recurrence, fused device dots"""
    return len(pre_ops) == 1 and tr is tr_cg and hasattr(dot_op, 'fused')


def cd_solve_dev(x, b, fwd_op, pre_ops, dot_op, criterion, roundoff=25):
    """cd_solve (reference cd_solve.py:35-107 with one preconditioner and tr_cg) for the TOP level of a chain: the same
    iterates, with the step lengths formed on the device by the fused dot kernels.  The only host synchronisation left in
    an iteration is the one the convergence monitor needs (`criterion` evaluates |r|^2 / d0 and logs it); the reference
    loop synchronises four times (delta, d.Ad, the orthogonalisation dot and the monitor)."""
    (pre_op,) = pre_ops
    residual = b - fwd_op(x)
    d = pre_op(residual)
    it = 0
    while not criterion(it, x, residual):
        Ad = fwd_op(d)
        delta = dot_op.fused(d, residual)
        t3 = dot_op.fused(d, Ad, num=delta[0:1])                 # [d.Ad, alpha, -alpha]
        it += 1
        if it % roundoff == 0:
            x = _axpy(x, t3[1:2], d)
            residual = b - fwd_op(x)
        else:
            _update_pair(x, d, residual, Ad, t3[1:2])
        dn = pre_op(residual)
        d = _axpy(dn, dot_op.fused(dn, Ad, den=t3[0:1], scale=-1.0)[1:2], d)
    return it


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
