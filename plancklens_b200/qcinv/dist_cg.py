"""Masked-sky CG filters with the full-resolution forward operator split by m over the GPUs of one box
("CG dots become a scalar all-reduce; alm stay m-distributed between calls").

No counterpart in the reference, where one filter is one process (`qcinv/multigrid.py:45-69`, `cd_solve.py:35-107`); the
solver below is that same preconditioned conjugate-gradient recurrence -- same operators, same chain description, same
monitor -- with its vectors laid out over the ranks of a `torch.distributed` (NCCL) process group:

* `x` (solution), `d` (search direction) and the preconditioned residual are REPLICATED (identical on every rank);
* `b`, the residual `r` and `A d` are M-DISTRIBUTED: a rank holds the rows of its own m (`dist_sht.partition`), zeros elsewhere.

One iteration per rank: the forward operator on its share -- Legendre synthesis of its m columns for all rings (peer stores
over NVLink), ring FFTs + N^-1 on its own rings, analysis of its own m (`dist_sht.DistPlan`, S^-1 x folded into the
analysis output) -- three dot products as local partial sums + one three-element all-reduce each, one all-reduce of the
residual (33 MB at lmax 2048) to feed the multigrid preconditioner, which every rank applies in full (its coarse levels
are launch-latency bound and would not speed up by splitting).  Scaling is therefore bounded by the replicated
preconditioner (Amdahl): what is split is the two full-resolution transforms per iteration.

Every rank takes the same branches: the convergence test uses the all-reduced residual norm.
"""
import numpy as np
import torch

from .. import dist_sht, sht
from . import cd_monitors, opfilt_pp, opfilt_tt, util
from .util_alm import dalm, eblm


def _comps(v):
    return [v] if isinstance(v, dalm) else [v.elm, v.blm]


def _like(v, ts):
    return dalm(ts[0], v.lmax) if isinstance(v, dalm) else eblm([dalm(ts[0], v.lmax), dalm(ts[1], v.lmax)])


class dist_chain:
    """Runs `chain.solve` (a `multigrid.multigrid_chain` on `opfilt_tt` or `opfilt_pp`) with the top-level forward
    operator m-partitioned.  Build it on every rank of the group after the chain exists; call `solve` collectively."""

    def __init__(self, chain, group=None):
        import torch.distributed as dist
        assert dist.is_initialized(), "dist_chain needs an initialised torch.distributed process group (torchrun)"
        self.dist, self.group = dist, group
        self.chain = chain
        self.opfilt = chain.opfilt
        self.pol = chain.opfilt is opfilt_pp
        assert self.pol or chain.opfilt is opfilt_tt, "m-distributed CG covers opfilt_tt and opfilt_pp"
        nf = chain.n_inv_filt
        self.nf = nf
        self.lmax = len(nf.b_transf) - 1
        self.plan = dist_sht.DistPlan(nf.nside, self.lmax, group=group)
        npix = 12 * nf.nside ** 2
        own = torch.zeros(npix, dtype=torch.float64, device='cuda')
        for lo, hi in self.plan.pixel_ranges():
            own[lo:hi] = 1.0
        if self.pol:
            nf._load_ninv()
            assert len(nf.n_inv) == 1 and not nf.wmarg, "m-distributed P filter: one N^-1 map, no template marginalisation"
            self.ninv_own = nf._ninv_d[0] * own
        else:
            assert len(nf._tmaps) == 0, "m-distributed T filter: monopole / dipole marginalisation only"
            self.ninv_own = nf._ninv_d * own
        # work maps stay zero outside this rank's rings: the N^-1 kernels then see (and leave) zeros there
        self._maps = [torch.zeros(npix, dtype=torch.float64, device='cuda') for _ in range(2 if self.pol else 1)]
        self._fl = {}

    # ---- pieces
    def _allreduce(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t

    def _dot(self, a, b, num=None, den=None, scale=1.0):
        """[s, r, -r] with s = the GLOBAL dot product: local fused partial sum, all-reduce, then the step-length ratio"""
        part = self.opfilt.dot_op().fused(a, b)
        s = self._allreduce(part[0:1].clone())
        if num is not None:
            r = sht.scalar_ratio(num, s, scale)
        elif den is not None:
            r = sht.scalar_ratio(s, den, scale)
        else:
            return s, None, None
        return s, r, sht.scalar_ratio(r, torch.ones_like(r), -1.0)

    def fwd(self, x):
        """(A x) on this rank's m rows, zeros elsewhere; x replicated"""
        nf, lmax, plan = self.nf, self.lmax, self.plan
        if self.pol:
            q, u = plan.alm2map_spin(x.elm.t, x.blm.t, 2, flg=nf._fl('ein', lmax), flc=nf._fl('bin', lmax), out=(self._maps[0], self._maps[1]))
            sht.map_mul2(q, u, self.ninv_own)
            if 'sl' not in self._fl:
                sl = opfilt_pp.alm_filter_sinv(self.chain.s_cls, lmax).slinv
                assert not np.any(sl[:, 0, 1]), "m-distributed P filter: diagonal S^-1 only"
                self._fl['sl'] = (sht.dev_fl(sl[:, 0, 0], lmax), sht.dev_fl(sl[:, 1, 1], lmax))
            se, sb = self._fl['sl']
            e, b = plan.map2alm_spin(q, u, 2, flg=nf._fl('eout', lmax), flc=nf._fl('bout', lmax), reduce=False,
                                     add=(x.elm.t, se, x.blm.t, sb))
            return eblm([dalm(e, lmax), dalm(b, lmax)])
        t = plan.alm2map(x.t, fl=nf.fl_in(lmax), out=self._maps[0])
        self._ninv_t(t)
        if 'cli' not in self._fl:
            self._fl['cli'] = sht.dev_fl(opfilt_tt._cli(self.chain.s_cls['tt']), lmax)
        return dalm(plan.map2alm(t, fl=nf.fl_out(lmax), reduce=False, add=(x.t, self._fl['cli'])), lmax)

    def _ninv_t(self, t):
        """opfilt_tt.apply_map on this rank's pixels: the template sums are all-reduced over ranks"""
        nf = self.nf
        p0 = sht.get_plan(nf.nside, nf._plan0.lmax)
        if len(nf.templates) != 0:
            p0.modes_dot(t, w=self.ninv_own, out=nf._sums)
            self._allreduce(nf._sums)
            p0.modes_sub(t, self.ninv_own, nf._sums, nf._pinv_d)
        else:
            sht.map_mul(t, self.ninv_own)

    def prep(self, maps):
        """b = B^t N^-1 d on this rank's m rows (calc_prep of the reference operators); `maps`: the full data map(s)"""
        nf, lmax, plan = self.nf, self.lmax, self.plan
        if self.pol:
            q = self._maps[0].copy_(sht.dev_map(maps[0]))
            u = self._maps[1].copy_(sht.dev_map(maps[1]))
            sht.map_mul2(q, u, self.ninv_own)
            e, b = plan.map2alm_spin(q, u, 2, flg=nf._fl('eout', lmax), flc=nf._fl('bout', lmax), reduce=False)
            return eblm([dalm(e, lmax), dalm(b, lmax)])
        t = self._maps[0].copy_(sht.dev_map(maps))
        self._ninv_t(t)
        return dalm(plan.map2alm(t, fl=nf.fl_out(lmax), reduce=False), lmax)

    def _replicate(self, v):
        """sum of the m-distributed rows of all ranks = the full vector, on every rank"""
        ts = [c.t.clone() for c in _comps(v)]
        for t in ts:
            self._allreduce(torch.view_as_real(t))
        return _like(v, ts)

    # ---- the solver
    def solve(self, soltn, maps, roundoff=25):
        """soltn (replicated device vector, `dalm` or `eblm` of `dalm`, zero or warm start) <- inverse-variance filtered
        solution of the data map(s); collective.  Returns the top-level iteration count (reference: multigrid.py:45-69)."""
        chain = self.chain
        (pre_op,) = chain.bstage.pre_ops
        b = self.prep(maps)
        d0 = float(self._dot(b, b)[0].item())
        logger = (lambda it, eps, stage=chain.bstage, **kw: chain.log(stage, it, eps, **kw))
        chain.watch = util.stopwatch()
        chain.iter_tot = 0
        glob_dot = lambda a, bb: float(self._dot(a, bb)[0].item())
        monitor = cd_monitors.monitor_basic(glob_dot, logger=logger, iter_max=chain.bstage.iter_max,
                                            eps_min=chain.bstage.eps_min, d0=d0)
        x = soltn
        r = b.copy() if x.is_zero() else b - self.fwd(x)
        d = pre_op(self._replicate(r))
        it = 0
        while not monitor(it, x, r):
            Ad = self.fwd(d)
            delta = self._dot(d, r)[0]
            dTAd, alpha, malpha = self._dot(d, Ad, num=delta)
            it += 1
            for cx, cd in zip(_comps(x), _comps(d)):
                sht.alm_axpy(cx.t, cd.t, alpha)
                cx.zero = False
            if it % roundoff == 0:
                r = b - self.fwd(x)
            else:
                for cr, ca in zip(_comps(r), _comps(Ad)):
                    sht.alm_axpy(cr.t, ca.t, malpha)
            dn = pre_op(self._replicate(r))
            _, beta, _ = self._dot(dn, Ad, den=dTAd, scale=-1.0)
            for cn, cd in zip(_comps(dn), _comps(d)):
                sht.alm_axpy(cn.t, cd.t, beta)
            d = dn
        getattr(self.opfilt, 'apply_fini')(x, chain.s_cls, chain.n_inv_filt)
        chain.niter, chain.last_monitor = it, monitor
        return it
