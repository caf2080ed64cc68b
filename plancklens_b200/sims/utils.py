"""Simulation-library wrappers (reference: plancklens/sims/utils.py)."""
import numpy as np


class sim_lib_shuffle:
    """Remaps simulation indices; index -1 conventionally points at the data map."""

    def __init__(self, sim_lib, shuffle_dict):
        self.sim_lib = sim_lib
        self._shuffle = shuffle_dict

    def get_sim_tmap(self, idx):
        return self.sim_lib.get_sim_tmap(int(self._shuffle[idx]))

    def get_sim_pmap(self, idx):
        return self.sim_lib.get_sim_pmap(int(self._shuffle[idx]))

    def __getattr__(self, name):
        # device-resident accessors exist when the wrapped library has them (maps.cmb_maps.get_sim_*map_dev)
        if name in ('get_sim_tmap_dev', 'get_sim_pmap_dev') and hasattr(self.sim_lib, name):
            fun = getattr(self.sim_lib, name)
            return lambda idx: fun(int(self._shuffle[idx]))
        raise AttributeError(name)

    def hashdict(self):
        return {'sim_lib': self.sim_lib.hashdict(), 'shuffle': self._shuffle}


class _sim_lib_sum:
    """Weighted sum of the maps of several simulation libraries, applied to the indices `_applies(idx)` selects; the
    other indices get the first library alone, times its weight."""
    _tag = None

    def __init__(self, sim_libs, weights=None):
        self.sim_libs = sim_libs
        self.w = np.ones(len(sim_libs)) if weights is None else weights

    @staticmethod
    def _applies(idx):
        raise NotImplementedError

    def _libs(self, idx):
        n = len(self.sim_libs) if self._applies(idx) else 1
        return list(zip(self.sim_libs[:n], self.w[:n]))

    def get_sim_tmap(self, idx):
        return sum(s.get_sim_tmap(idx) * w for s, w in self._libs(idx))

    def get_sim_pmap(self, idx):
        q, u = 0., 0.
        for s, w in self._libs(idx):
            _q, _u = s.get_sim_pmap(idx)
            q, u = q + w * _q, u + w * _u
        return q, u

    def hashdict(self):
        ret = {'lib': self._tag}
        for i, (s, w) in enumerate(zip(self.sim_libs, self.w)):
            ret['sim_lib ' + str(i)] = s.hashdict()
            ret['w ' + str(i)] = w
        return ret


class sim_lib_add_sim(_sim_lib_sum):
    """Sum of simulation libraries for simulation indices (>= 0); the data (-1) comes from the first one alone
    (reference: sims/utils.py:19-53)."""
    _tag = 'add_sim'

    @staticmethod
    def _applies(idx):
        return idx >= 0


class sim_lib_add_dat(_sim_lib_sum):
    """Sum of simulation libraries for the data index (< 0) only (reference: sims/utils.py:57-95)."""
    _tag = 'add_dat'

    @staticmethod
    def _applies(idx):
        return idx < 0
