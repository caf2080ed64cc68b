"""Small helpers of the CG package (reference: plancklens/qcinv/util.py)."""
import time

import numpy as np

from .. import hp, utils


class dt:
    def __init__(self, _dt):
        self.dt = _dt

    def __str__(self):
        return '%02d:%02d:%02d' % (self.dt // 3600, (self.dt % 3600) // 60, self.dt % 60)

    def __int__(self):
        return int(self.dt)


class stopwatch:
    def __init__(self):
        self.st = time.time()
        self.lt = self.st

    def lap(self):
        now = time.time()
        ret = (dt(now - self.st), dt(now - self.lt))
        self.lt = now
        return ret

    def elapsed(self):
        now = time.time()
        self.lt = now
        return dt(now - self.st)


class jit:
    """Instantiate-on-first-use wrapper (reference: qcinv/util.py:39-61)."""

    def __init__(self, ctype, *cargs, **ckwds):
        self.__dict__['_jit_spec'] = (ctype, cargs, ckwds)
        self.__dict__['_jit_obj'] = None

    def instantiate(self):
        ctype, cargs, ckwds = self.__dict__['_jit_spec']
        self.__dict__['_jit_obj'] = ctype(*cargs, **ckwds)

    def __getattr__(self, attr):
        if self.__dict__['_jit_obj'] is None:
            self.instantiate()
        return getattr(self.__dict__['_jit_obj'], attr)

    def __setattr__(self, attr, val):
        if self.__dict__['_jit_obj'] is None:
            self.instantiate()
        setattr(self.__dict__['_jit_obj'], attr, val)


def read_map(m):
    """Map given as array, callable, path (optionally 'path,field') or list of those to multiply
    (reference: qcinv/util.py:63-79)."""
    if callable(m):
        return m()
    if isinstance(m, list):
        out = read_map(m[0])
        for m2 in m[1:]:
            out = out * read_map(m2)
        return out
    if not isinstance(m, str):
        return m
    if ',' not in m:
        return hp.read_map(m)
    fname, field = m.split(',')
    return hp.read_map(fname, field=int(field))


load_map = read_map


def mask_hash(m, dtype=bool):
    if m is None:
        return "none"
    if isinstance(m, list):
        return ''.join(mask_hash(x, dtype=dtype) for x in m)
    if isinstance(m, str):
        return m.replace('/', '_sl_').replace('.', '_')
    if isinstance(m, np.ndarray):
        return utils.clhash(m, dtype=dtype)
    if callable(m):
        return 'callable'
    raise AssertionError('not implemented')
