"""Wrappers around filtering libraries (reference: plancklens/filt/filt_util.py:39-236)."""
import numpy as np

from .. import hp, utils


class library_ftl:
    """Rescales the filtered alms of another library by isotropic functions (lmin cuts, ...)
    (reference: filt_util.py:39-183)."""

    def __init__(self, ivfs, lmax, lfilt_t, lfilt_e, lfilt_b):
        assert len(lfilt_t) > lmax and len(lfilt_e) > lmax and len(lfilt_b) > lmax
        self.ivfs = ivfs
        self.lmax = lmax
        self.lfilt_t, self.lfilt_e, self.lfilt_b = lfilt_t, lfilt_e, lfilt_b
        self.lib_dir = ivfs.lib_dir

    def hashdict(self):
        return {'ivfs': self.ivfs.hashdict(), 'filt_t': utils.clhash(self.lfilt_t[:self.lmax + 1]),
                'filt_e': utils.clhash(self.lfilt_e[:self.lmax + 1]), 'filt_b': utils.clhash(self.lfilt_b[:self.lmax + 1])}

    def get_fmask(self):
        return self.ivfs.get_fmask()

    def get_tal(self, a):
        return self.ivfs.get_tal(a)

    def get_ftl(self):
        return self.ivfs.get_ftl()[:self.lmax + 1] * self.lfilt_t[:self.lmax + 1]

    def get_fel(self):
        return self.ivfs.get_fel()[:self.lmax + 1] * self.lfilt_e[:self.lmax + 1]

    def get_fbl(self):
        return self.ivfs.get_fbl()[:self.lmax + 1] * self.lfilt_b[:self.lmax + 1]

    def _cut(self, alm, fl):
        """alm brought to self.lmax (truncated or zero-padded, reference: filt_util.py:10-37) times fl"""
        lmax_in = hp.Alm.getlmax(alm.size)
        if lmax_in >= self.lmax:
            out = utils.alm_copy(alm, lmax=self.lmax)
        else:
            out = np.zeros(hp.Alm.getsize(self.lmax), dtype=complex)
            for m in range(lmax_in + 1):
                i = m * (2 * lmax_in + 1 - m) // 2 + m
                o = m * (2 * self.lmax + 1 - m) // 2 + m
                out[o:o + lmax_in + 1 - m] = alm[i:i + lmax_in + 1 - m]
        return hp.almxfl(out, fl)

    def get_sim_tlm(self, idx):
        return self._cut(self.ivfs.get_sim_tlm(idx), self.lfilt_t)

    def get_sim_elm(self, idx):
        return self._cut(self.ivfs.get_sim_elm(idx), self.lfilt_e)

    def get_sim_blm(self, idx):
        return self._cut(self.ivfs.get_sim_blm(idx), self.lfilt_b)

    def get_sim_tmliklm(self, idx):
        return self._cut(self.ivfs.get_sim_tmliklm(idx), self.lfilt_t)

    def get_sim_emliklm(self, idx):
        return self._cut(self.ivfs.get_sim_emliklm(idx), self.lfilt_e)

    def get_sim_bmliklm(self, idx):
        return self._cut(self.ivfs.get_sim_bmliklm(idx), self.lfilt_b)


class library_shuffle:
    """Filtering library with remapped simulation indices (reference: filt_util.py:186-236)."""

    def __init__(self, ivfs, idxs):
        self.ivfs = ivfs
        self.idxs = idxs
        self.lib_dir = getattr(ivfs, 'lib_dir', None)

    def hashdict(self):
        return {'ivfs': self.ivfs.hashdict(), 'idxs': self.idxs}

    def get_fmask(self):
        return self.ivfs.get_fmask()

    def get_tal(self, a):
        return self.ivfs.get_tal(a)

    def get_ftl(self):
        return self.ivfs.get_ftl()

    def get_fel(self):
        return self.ivfs.get_fel()

    def get_fbl(self):
        return self.ivfs.get_fbl()

    def get_sim_tlm(self, idx):
        return self.ivfs.get_sim_tlm(self.idxs[idx])

    def get_sim_elm(self, idx):
        return self.ivfs.get_sim_elm(self.idxs[idx])

    def get_sim_blm(self, idx):
        return self.ivfs.get_sim_blm(self.idxs[idx])

    def get_sim_tmliklm(self, idx):
        return self.ivfs.get_sim_tmliklm(self.idxs[idx])

    def get_sim_emliklm(self, idx):
        return self.ivfs.get_sim_emliklm(self.idxs[idx])

    def get_sim_bmliklm(self, idx):
        return self.ivfs.get_sim_bmliklm(self.idxs[idx])
