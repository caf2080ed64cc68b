"""Random-phase libraries (reference: plancklens/sims/phas.py:137-195).

The reference stores numpy RNG states in sqlite so that every phase can be regenerated; here each
(idx, field) phase is regenerated from a counter-based seed instead -- same interface (`get_sim`, `hashdict`,
`nfields`, `lmax` / `shape`), no database.  The draws follow the recipe at phas.py:162-168:
alm = (N(0,1) + i N(0,1)) / sqrt(2) with real N(0,1) at m = 0.
"""
import numpy as np


class lib_phas:
    """Unit-variance Gaussian alm phases, `nfields` independent fields up to lmax."""

    def __init__(self, lib_dir, nfields, lmax, seed=10000):
        self.lib_dir = lib_dir
        self.nfields = nfields
        self.lmax = lmax
        self.seed = seed

    def _one(self, idx, idf):
        rng = np.random.default_rng([self.seed, int(idx) & 0xffffffff, idf])
        n = (self.lmax + 1) * (self.lmax + 2) // 2
        alm = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.)
        alm[:self.lmax + 1] = np.sqrt(2.) * alm[:self.lmax + 1].real     # m = 0 block comes first in the layout
        return alm

    def get_sim(self, idx, idf=None, phas_only=False):
        if idf is not None:
            assert idf < self.nfields, (idf, self.nfields)
            return self._one(idx, idf)
        return np.array([self._one(idx, i) for i in range(self.nfields)])

    def is_full(self):
        return True

    def hashdict(self):
        return {'nfields': self.nfields, 'lmax': self.lmax, 'seed': self.seed}


class pix_lib_phas:
    """Unit-variance Gaussian pixel-space phases (noise maps), `nfields` fields of a given shape."""

    def __init__(self, lib_dir, nfields, shape, seed=20000):
        self.lib_dir = lib_dir
        self.nfields = nfields
        self.shape = shape
        self.seed = seed

    def get_sim(self, idx, idf=None, phas_only=False):
        def one(i):
            return np.random.default_rng([self.seed, int(idx) & 0xffffffff, i]).standard_normal(self.shape)
        if idf is not None:
            assert idf < self.nfields, (idf, self.nfields)
            return one(idf)
        return np.array([one(i) for i in range(self.nfields)])

    def is_full(self):
        return True

    def hashdict(self):
        return {'nfields': self.nfields, 'shape': self.shape, 'seed': self.seed}
