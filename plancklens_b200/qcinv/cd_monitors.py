"""Convergence monitor of the CG solver (reference: plancklens/qcinv/cd_monitors.py:29-41).

Same plug-in surface -- `monitor_basic(dot_op, iter_max, eps_min, logger, d0)` called as
`criterion(iter, soltn, resid) -> bool` by `cd_solve` -- with the residual history kept in `trace`, which is what
the parity tests compare against the reference's log.
"""
import sys

import numpy as np

from . import util


def logger_basic(iter, eps, watch=None, **kwargs):
    sys.stdout.write('[%s] %s\n' % (watch.elapsed(), (iter, eps)))


def logger_none(iter, eps, watch=None, **kwargs):
    return 0


class monitor_basic:
    """Stop rule `iter >= iter_max or |r|^2 <= eps_min^2 d0`; d0 is the squared norm seen at iteration 0 unless given.

    The squared residual norm costs one `dot_op` (a device reduction + one host read) per call; the fixed-iteration
    multigrid stages bypass the monitor altogether (`cd_solve.cd_solve_fixed`)."""

    def __init__(self, dot_op, iter_max=1000, eps_min=1.0e-10, logger=logger_basic, d0=None):
        self.dot_op, self.logger = dot_op, logger
        self.iter_max, self.eps_min = iter_max, eps_min
        self.d0 = d0
        self.trace = []                 # (iter, eps) of every call
        self.watch = util.stopwatch()

    def criterion(self, iter, soltn, resid):
        norm2 = self.dot_op(resid, resid)
        if self.d0 is None and iter == 0:
            self.d0 = norm2
        eps = float(np.sqrt(norm2 / self.d0))
        self.trace.append((iter, eps))
        if self.logger is not None:
            self.logger(iter, eps, watch=self.watch, soltn=soltn, resid=resid)
        converged = norm2 <= self.d0 * self.eps_min ** 2
        return bool(converged or iter >= self.iter_max)

    __call__ = criterion
