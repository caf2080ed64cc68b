"""Multigrid-preconditioned CG chain (reference: plancklens/qcinv/multigrid.py).

Same `chain_descr` rows `[id, pre_ops, lmax, nside, iter_max, eps_min, tr, cache]` and the same preconditioner
mini-language -- `split(low, lsplit, high)`, `diag_cl`, `dense`, `dense(file)`, `stage(k)` -- as the reference
(multigrid.py:37, :113-160).  All vectors stay on the GPU from `calc_prep` to `apply_fini`.
"""
import copy
import gc
import os
import re
import sys
import threading

import torch

from . import cd_monitors, cd_solve, util, util_alm


class multigrid_stage(object):
    def __init__(self, ids, pre_ops_descr, lmax, nside, iter_max, eps_min, tr, cache):
        self.depth = ids
        self.pre_ops_descr = pre_ops_descr
        self.lmax = lmax
        self.nside = nside
        self.iter_max = iter_max
        self.eps_min = eps_min
        self.tr = tr
        self.cache = cache
        self.pre_ops = []


class multigrid_chain:
    def __init__(self, opfilt, chain_descr, s_cls, n_inv_filt, debug_log_prefix=None, plogdepth=0):
        self.debug_log_prefix = debug_log_prefix
        self.plogdepth = plogdepth
        self.opfilt = opfilt
        self.chain_descr = chain_descr
        self.s_cls = s_cls
        self.n_inv_filt = n_inv_filt
        self.iter_tot = 0
        self.last_monitor = None

        stages = {}
        for [sid, pre_ops_descr, lmax, nside, iter_max, eps_min, tr, cache] in self.chain_descr:
            stages[sid] = multigrid_stage(sid, pre_ops_descr, lmax, nside, iter_max, eps_min, tr, cache)
            for descr in pre_ops_descr:   # deeper stages were built by earlier rows
                stages[sid].pre_ops.append(parse_pre_op_descr(descr, opfilt=self.opfilt, s_cls=self.s_cls,
                                                               n_inv_filt=self.n_inv_filt, stages=stages,
                                                               lmax=lmax, nside=nside, chain=self))
        self.bstage = stages[0]
        # B200: the fixed-iteration inner stages are a static launch sequence (cd_solve_fixed); replay them as one
        # CUDA graph per preconditioner call instead of a few thousand ctypes launches (PLK_CG_GRAPH=0 disables)
        if os.environ.get('PLK_CG_GRAPH', '1') != '0':
            self.bstage.pre_ops = [graphed_op(op) if (_has_stage(op) and _graph_safe(op)) else op
                                   for op in self.bstage.pre_ops]

    def solve(self, soltn, tpn_map, apply_fini='', dot_op=None):
        """soltn <- inverse-variance filtered solution for the data map(s) tpn_map.

        soltn may be a device vector (`dalm` / `eblm` of `dalm`) or the numpy containers of the reference, in
        which case it is updated in place on return (reference: multigrid.py:45-69)."""
        finifunc = getattr(self.opfilt, 'apply_fini%s' % apply_fini)
        self.watch = util.stopwatch()
        self.iter_tot = 0
        self.prev_eps = None
        if dot_op is None:
            dot_op = self.opfilt.dot_op()
        logger = (lambda iter, eps, stage=self.bstage, **kwargs: self.log(stage, iter, eps, **kwargs))

        dsol, writeback = _to_device(soltn)
        tpn_alm = self.opfilt.calc_prep(tpn_map, self.s_cls, self.n_inv_filt)
        monitor = cd_monitors.monitor_basic(dot_op, logger=logger, iter_max=self.bstage.iter_max,
                                            eps_min=self.bstage.eps_min, d0=dot_op(tpn_alm, tpn_alm))
        fwd_op = self.opfilt.fwd_op(self.s_cls, self.n_inv_filt)
        if cd_solve.can_solve_dev(self.bstage.pre_ops, dot_op, self.bstage.tr) and os.environ.get('PLK_CG_DEVTOP', '1') != '0':
            # step lengths stay on the device: one host synchronisation per iteration (the monitor's) instead of four
            self.niter = cd_solve.cd_solve_dev(dsol, tpn_alm, fwd_op, self.bstage.pre_ops, dot_op, monitor)
        else:
            self.niter = cd_solve.cd_solve(dsol, tpn_alm, fwd_op, self.bstage.pre_ops, dot_op, monitor,
                                           tr=self.bstage.tr, cache=self.bstage.cache)
        finifunc(dsol, self.s_cls, self.n_inv_filt)
        self.last_monitor = monitor
        if writeback is not None:
            writeback(dsol)
        return dsol

    def log(self, stage, iter, eps, **kwargs):
        self.iter_tot += 1
        elapsed = self.watch.elapsed()
        if stage.depth > self.plogdepth:
            return
        log_str = '   ' * stage.depth + '(%4d, %04d) [%s] (%d, %.8f)' % (stage.nside, stage.lmax, str(elapsed), iter, eps) + '\n'
        sys.stdout.write(log_str)
        if self.debug_log_prefix is not None:
            with open(self.debug_log_prefix + 'stage_all.dat', 'a') as log:
                log.write(log_str)
            with open(self.debug_log_prefix + 'stage_' + str(stage.depth) + '.dat', 'a') as log:
                log.write('%05d %05d %10.6e %05d %s\n' % (self.iter_tot, int(elapsed), eps, iter, str(elapsed)))


def _has_stage(op):
    if isinstance(op, pre_op_split):
        return _has_stage(op.pre_op_low) or _has_stage(op.pre_op_hgh)
    return isinstance(op, pre_op_multigrid)


def _graph_safe(op):
    """True when applying `op` never synchronises with the host (so it can be stream-captured)."""
    if isinstance(op, pre_op_split):
        return _graph_safe(op.pre_op_low) and _graph_safe(op.pre_op_hgh)
    if isinstance(op, pre_op_multigrid):
        return op.fixed and all(_graph_safe(p) for p in op.pre_ops)
    return type(op).__name__ in ('pre_op_diag', 'pre_op_dense_tt', 'pre_op_dense_pp', 'pre_op_dense_tp')


def _vec_tensors(v):
    if isinstance(v, util_alm.dalm):
        return [v.t]
    if isinstance(v, util_alm.teblm):
        return [v.tlm.t, v.elm.t, v.blm.t]
    return [v.elm.t, v.blm.t]


def _vec_clone(v):
    if isinstance(v, util_alm.dalm):
        return util_alm.dalm(v.t.clone(), v.lmax, v.zero)
    if isinstance(v, util_alm.teblm):
        return util_alm.teblm([_vec_clone(v.tlm), _vec_clone(v.elm), _vec_clone(v.blm)])
    return util_alm.eblm([_vec_clone(v.elm), _vec_clone(v.blm)])


_CAPTURE_LOCK = threading.Lock()      # torch.cuda.graph synchronises and trims the allocator on entry: one capture at a time


class graphed_op:
    """Applies a host-synchronisation-free preconditioner through a CUDA graph.

    Call 1 runs eagerly (plans, tables and per-l factors get allocated and uploaded); call 2 is captured
    (torch.cuda.graph on the launching stream: every kernel is a libplk_b200 launch); later calls copy the
    argument into the captured input, replay, and return a copy of the captured output."""

    def __init__(self, op):
        self.op = op
        self.calls = 0
        self.graph = None
        self.v_in = self.v_out = None
        # The preconditioner is a chain of small dependent kernels.  When another lane's full-resolution transforms are in
        # flight (filt_simple.library_sepTP runs the T and the P filter side by side) it goes through a high-priority
        # stream: the block scheduler then places its few blocks ahead of the pending blocks of the big grid instead of
        # behind all of them.  PLK_CG_PRIO=0 replays on the calling stream.
        self.prio = os.environ.get('PLK_CG_PRIO', '1') != '0'
        self.stream = None

    def __call__(self, v):
        return self.calc(v)

    def calc(self, v):
        self.calls += 1
        if self.calls == 1:
            return self.op(v)
        if self.graph is None:
            self.v_in = _vec_clone(v)
            g = torch.cuda.CUDAGraph()
            # Dead reference cycles may own CUDA graphs / plans whose destructors call cudaFree, which is illegal while
            # a stream is capturing: collect them now and keep the cyclic collector off during the capture.
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            # own capture stream: torch's default one is shared by every capture of the process, and two lanes may be
            # capturing at the same time
            # lane 0 (temperature, the longer solve) outranks lane 1 (polarization); within a lane the preconditioner
            # outranks the full-resolution operator of the other lane's level
            from .. import sht
            self.stream = torch.cuda.Stream(priority=(-3 if sht.lane() == 0 else -1) if self.prio else 0)
            try:
                with _CAPTURE_LOCK:
                    with torch.cuda.graph(g, stream=self.stream, capture_error_mode='thread_local'):
                        self.v_out = self.op(self.v_in)
            finally:
                if gc_was_on:
                    gc.enable()
            self.graph = g
        if not self.prio:
            for dst, src in zip(_vec_tensors(self.v_in), _vec_tensors(v)):
                dst.copy_(src)
            self.graph.replay()
            return _vec_clone(self.v_out)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for dst, src in zip(_vec_tensors(self.v_in), _vec_tensors(v)):
                dst.copy_(src)
            self.graph.replay()
        cur.wait_stream(self.stream)
        return _vec_clone(self.v_out)


def _to_device(soltn):
    """-> (device vector, writeback or None)"""
    if isinstance(soltn, util_alm.dalm):
        return soltn, None
    if isinstance(soltn, util_alm.teblm):
        if isinstance(soltn.tlm, util_alm.dalm):
            return soltn, None
        d = util_alm.teblm([util_alm.dalm.from_numpy(c) for c in (soltn.tlm, soltn.elm, soltn.blm)])

        def wb(v):
            t, e, b = v.numpy()
            soltn.tlm[:] = t
            soltn.elm[:] = e
            soltn.blm[:] = b
        return d, wb
    if isinstance(soltn, util_alm.eblm):
        if isinstance(soltn.elm, util_alm.dalm):
            return soltn, None
        d = util_alm.eblm([util_alm.dalm.from_numpy(soltn.elm), util_alm.dalm.from_numpy(soltn.blm)])

        def wb(v):
            e, b = v.numpy()
            soltn.elm[:] = e
            soltn.blm[:] = b
        return d, wb
    d = util_alm.dalm.from_numpy(soltn)

    def wb(v):
        soltn[:] = v.numpy()
    return d, wb


def parse_pre_op_descr(pre_op_descr, **kwargs):
    """Builds a preconditioner from its description string (reference: multigrid.py:113-160)."""
    m = re.match(r"split\((.*),\s*(.*),\s*(.*)\)\Z", pre_op_descr)
    if m:
        low_descr, lsplit, hgh_descr = m.groups()
        print('creating split preconditioner ', (low_descr, lsplit, hgh_descr))
        lsplit = int(lsplit)
        kw_low = copy.copy(kwargs)
        kw_low['lmax'] = lsplit
        kw_hgh = copy.copy(kwargs)
        kw_hgh['lmin'] = lsplit + 1
        return pre_op_split(lsplit, kwargs['lmax'], parse_pre_op_descr(low_descr, **kw_low),
                            parse_pre_op_descr(hgh_descr, **kw_hgh))
    if re.match(r"diag_cl\Z", pre_op_descr):
        return kwargs['opfilt'].pre_op_diag(kwargs['s_cls'], kwargs['n_inv_filt'])
    m = re.match(r"dense(\((.*)\))?\Z", pre_op_descr)
    if m:
        fname = m.group(2)
        if fname == '' or fname is None:
            fname = None
        print('creating dense preconditioner. (nside = %d, lmax = %d, cache = %s)' % (kwargs['nside'], kwargs['lmax'], fname))
        fwd_op = kwargs['opfilt'].fwd_op(kwargs['s_cls'], kwargs['n_inv_filt'].degrade(kwargs['nside']))
        return kwargs['opfilt'].pre_op_dense(kwargs['lmax'], fwd_op, cache_fname=fname)
    m = re.match(r"stage\((.*)\)\Z", pre_op_descr)
    if m:
        (stage_id,) = m.groups()
        print('creating multigrid preconditioner: stage_id = ', stage_id)
        stage = kwargs['stages'][int(stage_id)]
        logger = (lambda iter, eps, stage=stage, chain=kwargs['chain'], **kw: chain.log(stage, iter, eps, **kw))
        assert stage.lmax == kwargs['lmax']
        return pre_op_multigrid(kwargs['opfilt'], stage.lmax, stage.nside, kwargs['s_cls'],
                                kwargs['n_inv_filt'].degrade(stage.nside), stage.pre_ops, logger, stage.tr,
                                stage.cache, stage.iter_max, stage.eps_min)
    assert 0, 'pre_op_descr' + pre_op_descr + ' is unrecognized!'


class pre_op_split:
    """Low-l block through one operator, high-l through another, spliced at lsplit (reference: multigrid.py:163-182)."""

    def __init__(self, lsplit, lmax, pre_op_low, pre_op_hgh):
        self.lsplit = lsplit
        self.lmax = lmax
        self.pre_op_low = pre_op_low
        self.pre_op_hgh = pre_op_hgh
        self.iter = 0

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, talm):
        self.iter += 1
        filt = getattr(self.pre_op_hgh, '_filt_d', None)
        if isinstance(talm, util_alm.dalm) and filt is not None and talm.lmax == self.lmax:
            # device vector, `diag_cl` above lsplit (opfilt_tt.pre_op_diag): the high-l almxfl and the splice are one
            # kernel, and the operators read their argument without the defensive copies (none of them modifies it)
            low = self.pre_op_low.calc_low(talm, self.lsplit) if hasattr(self.pre_op_low, 'calc_low') \
                else self.pre_op_low(talm if talm.lmax == self.lsplit else util_alm.alm_copy(talm, lmax=self.lsplit))
            from .. import sht
            return util_alm.dalm(sht.alm_splice_xfl(low.t, talm.t, filt, self.lsplit), self.lmax)
        talm_low = self.pre_op_low(util_alm.alm_copy(talm, lmax=self.lsplit))
        talm_hgh = self.pre_op_hgh(util_alm.alm_copy(talm, lmax=self.lmax))
        return util_alm.alm_splice(talm_low, talm_hgh, self.lsplit)


class pre_op_multigrid:
    """A few CG iterations on a degraded grid used as a preconditioner (reference: multigrid.py:185-215)."""

    def __init__(self, opfilt, lmax, nside, s_cls, n_inv_filt, pre_ops, logger, tr, cache, iter_max, eps_min):
        self.opfilt = opfilt
        self.fwd_op = opfilt.fwd_op(s_cls, n_inv_filt)
        self.lmax = lmax
        self.nside = nside
        self.s_cls = s_cls
        self.pre_ops = pre_ops
        self.logger = logger
        self.tr = tr
        self.cache = cache
        self.iter_max = iter_max
        self.eps_min = eps_min
        self.fixed = cd_solve.can_solve_fixed(pre_ops, opfilt.dot_op(), tr, iter_max, eps_min) \
            and os.environ.get('PLK_CG_FIXED', '1') != '0'

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, talm):
        if self.fixed:
            soltn = talm * 0.0
            if getattr(talm, 'lmax', None) == self.lmax:
                # the stage works at the lmax of its argument (every row of the default chains): no truncating copy
                # before, no splice after -- cd_solve_fixed copies its right-hand side itself and never writes to it
                cd_solve.cd_solve_fixed(soltn, talm, self.fwd_op, self.pre_ops, self.opfilt.dot_op(), self.iter_max)
                return soltn
            cd_solve.cd_solve_fixed(soltn, util_alm.alm_copy(talm, lmax=self.lmax), self.fwd_op, self.pre_ops,
                                    self.opfilt.dot_op(), self.iter_max)
            return util_alm.alm_splice(soltn, talm, self.lmax)
        monitor = cd_monitors.monitor_basic(self.opfilt.dot_op(), iter_max=self.iter_max, eps_min=self.eps_min,
                                            logger=self.logger)
        soltn = talm * 0.0
        cd_solve.cd_solve(soltn, util_alm.alm_copy(talm, lmax=self.lmax), self.fwd_op, self.pre_ops,
                          self.opfilt.dot_op(), monitor, tr=self.tr, cache=self.cache)
        return util_alm.alm_splice(soltn, talm, self.lmax)
