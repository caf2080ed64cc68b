"""Simulation-library wrappers (reference: plancklens/sims/utils.py:3-16)."""


class sim_lib_shuffle:
    """Remaps simulation indices; index -1 conventionally points at the data map."""

    def __init__(self, sim_lib, shuffle_dict):
        self.sim_lib = sim_lib
        self._shuffle = shuffle_dict

    def get_sim_tmap(self, idx):
        return self.sim_lib.get_sim_tmap(int(self._shuffle[idx]))

    def get_sim_pmap(self, idx):
        return self.sim_lib.get_sim_pmap(int(self._shuffle[idx]))

    def hashdict(self):
        return {'sim_lib': self.sim_lib.hashdict(), 'shuffle': self._shuffle}
