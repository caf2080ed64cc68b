"""Small object caches keyed by a file-name stem (reference: plancklens/helpers/cachers.py): none, memory, .npy, pickle."""
import os
import pickle as pk

import numpy as np


class cacher(object):
    def cache(self, fn, obj):
        assert 0, 'subclass this'

    def load(self, fn):
        assert 0, 'subclass this'

    def is_cached(self, fn):
        assert 0, 'subclass this'


class cacher_none(cacher):
    def cache(self, fn, obj):
        pass

    def is_cached(self, fn):
        return False


class cacher_mem(cacher):
    def __init__(self):
        self._cache = {}

    def cache(self, fn, obj):
        self._cache[fn] = np.copy(obj)

    def load(self, fn):
        assert fn in self._cache, fn
        return np.copy(self._cache[fn])

    def is_cached(self, fn):
        return fn in self._cache


class _cacher_file(cacher):
    ext = None

    def __init__(self, lib_dir, verbose=False):
        os.makedirs(lib_dir, exist_ok=True)
        self.lib_dir, self.verbose = lib_dir, verbose

    def _path(self, fn):
        assert self.ext not in fn, fn
        return os.path.join(self.lib_dir, fn + self.ext)

    def is_cached(self, fn):
        return os.path.exists(self._path(fn))

    def cache(self, fn, obj):
        self._write(self._path(fn), obj)
        if self.verbose:
            print("Cached " + fn + self.ext)

    def load(self, fn):
        p = self._path(fn)
        assert os.path.exists(p), p
        if self.verbose:
            print("Loading " + fn + self.ext)
        return self._read(p)


class cacher_npy(_cacher_file):
    ext = '.npy'
    _write = staticmethod(lambda p, obj: np.save(p, obj))
    _read = staticmethod(np.load)


class cacher_pk(_cacher_file):
    ext = '.pk'

    @staticmethod
    def _write(p, obj):
        with open(p, 'wb') as f:
            pk.dump(obj, f)

    @staticmethod
    def _read(p):
        with open(p, 'rb') as f:
            return pk.load(f)
