"""Convergence monitor of the CG solver (reference: plancklens/qcinv/cd_monitors.py:29-41)."""
import sys

import numpy as np

from . import util

logger_basic = (lambda iter, eps, watch=None, **kwargs:
                sys.stdout.write('[' + str(watch.elapsed()) + '] ' + str((iter, eps)) + '\n'))
logger_none = (lambda iter, eps, watch=None, **kwargs: 0)


class monitor_basic:
    """Stops when iter >= iter_max or |r|^2 <= eps_min^2 d0, d0 the norm at the first call (or given)."""

    def __init__(self, dot_op, iter_max=1000, eps_min=1.0e-10, logger=logger_basic, d0=None):
        self.dot_op = dot_op
        self.iter_max = iter_max
        self.eps_min = eps_min
        self.logger = logger
        self.d0 = d0
        self.watch = util.stopwatch()
        self.trace = []       # (iter, eps) history; what the parity tests compare against the reference

    def criterion(self, iter, soltn, resid):
        delta = self.dot_op(resid, resid)
        if iter == 0 and self.d0 is None:
            self.d0 = delta
        eps = np.sqrt(delta / self.d0)
        self.trace.append((iter, float(eps)))
        if self.logger is not None:
            self.logger(iter, eps, watch=self.watch, soltn=soltn, resid=resid)
        return bool(iter >= self.iter_max or delta <= self.eps_min ** 2 * self.d0)

    def __call__(self, *args):
        return self.criterion(*args)
