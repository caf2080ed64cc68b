"""Random-phase libraries (reference: plancklens/sims/phas.py:137-195).

The reference stores numpy RNG states in sqlite so that every phase can be regenerated; here each
(idx, field) phase is regenerated from a counter-based seed instead -- same interface (`get_sim`, `hashdict`,
`nfields`, `lmax` / `shape`), no database.  The draws follow the recipe at phas.py:162-168:
alm = (N(0,1) + i N(0,1)) / sqrt(2) with real N(0,1) at m = 0.

Two generators: numpy's (host, default) and, with `device=True`, the Philox4x32-10 kernels of libplk_b200
(`get_sim_dev` returns CUDA tensors; `get_sim` then returns the same numbers copied to the host, so a library is one
stream of randomness whichever accessor is used, and its hash says which generator it is).
"""
import numpy as np


def _stream_id(idx, idf):
    """Philox stream of (simulation index, field): idx = -1 (the data) maps to 0xffffffff"""
    return ((int(idx) & 0xffffffff) << 8) | (int(idf) & 0xff)


class lib_phas:
    """Unit-variance Gaussian alm phases, `nfields` independent fields up to lmax."""

    def __init__(self, lib_dir, nfields, lmax, seed=10000, device=False):
        self.lib_dir = lib_dir
        self.nfields = nfields
        self.lmax = lmax
        self.seed = seed
        self.device = device

    def get_sim_dev(self, idx, idf):
        """phase of field idf as a complex128 CUDA tensor (Philox; needs device=True so that host and device agree)"""
        assert self.device, "construct the library with device=True to draw on the GPU"
        assert idf < self.nfields, (idf, self.nfields)
        from .. import sht
        return sht.randn_alm(self.seed, _stream_id(idx, idf), self.lmax)

    def _one(self, idx, idf):
        if self.device:
            return self.get_sim_dev(idx, idf).cpu().numpy()
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
        ret = {'nfields': self.nfields, 'lmax': self.lmax, 'seed': self.seed}
        if self.device:
            ret['rng'] = 'philox4x32-10'
        return ret


class pix_lib_phas:
    """Unit-variance Gaussian pixel-space phases (noise maps), `nfields` fields of a given shape."""

    def __init__(self, lib_dir, nfields, shape, seed=20000, device=False):
        self.lib_dir = lib_dir
        self.nfields = nfields
        self.shape = shape
        self.seed = seed
        self.device = device

    def get_sim_dev(self, idx, idf, scale=1.0, add=None):
        """unit normals of field idf as a float64 CUDA tensor; `add + scale * phase` in one pass when `add` is given"""
        assert self.device, "construct the library with device=True to draw on the GPU"
        assert idf < self.nfields, (idf, self.nfields)
        from .. import sht
        n = int(np.prod(self.shape))
        return sht.randn(self.seed, _stream_id(idx, idf), n, scale=scale, add=add).reshape(self.shape)

    def get_sim(self, idx, idf=None, phas_only=False):
        def one(i):
            if self.device:
                return self.get_sim_dev(idx, i).cpu().numpy()
            return np.random.default_rng([self.seed, int(idx) & 0xffffffff, i]).standard_normal(self.shape)
        if idf is not None:
            assert idf < self.nfields, (idf, self.nfields)
            return one(idf)
        return np.array([one(i) for i in range(self.nfields)])

    def is_full(self):
        return True

    def hashdict(self):
        ret = {'nfields': self.nfields, 'shape': self.shape, 'seed': self.seed}
        if self.device:
            ret['rng'] = 'philox4x32-10'
        return ret
