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

    # device-resident accessors, present when the wrapped library has them (filt_simple.library_sepTP)
    def _cut_dev(self, alm, a):
        from .. import sht
        if not hasattr(self, '_lf_d'):
            self._lf_d = {}
        if a not in self._lf_d:
            self._lf_d[a] = sht.dev_fl({'t': self.lfilt_t, 'e': self.lfilt_e, 'b': self.lfilt_b}[a], self.lmax)
        if sht.alm_lmax(alm.numel()) != self.lmax:
            alm = sht.alm_copy(alm, self.lmax)          # truncates or zero-pads
        return sht.almxfl(alm, self._lf_d[a])

    def __getattr__(self, name):
        if name in ('get_sim_teblm_dev', 'get_sim_mliklm_dev') and hasattr(self.ivfs, name):
            inner = getattr(self.ivfs, name)
            return lambda idx, fields='teb': tuple(self._cut_dev(x, f) for f, x in zip(fields, inner(idx, fields)))
        if name in ('flush', 'prefetch_dev') and hasattr(self.ivfs, name):
            return getattr(self.ivfs, name)
        raise AttributeError(name)

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


class library_fml:
    """Rescales the filtered alms of another library by functions of m, alm -> f_m alm (reference: filt_util.py:106-182).

    The isotropic filters reported by get_ftl / get_fel / get_fbl carry the square root of the m-averaged weight
    (2 sum_{m <= l} f_m - f_0) / (2 l + 1)."""

    def __init__(self, ivfs, lmax, mfilt_t, mfilt_e, mfilt_b):
        assert len(mfilt_t) > lmax and len(mfilt_e) > lmax and len(mfilt_b) > lmax
        self.ivfs = ivfs
        self.lmax = lmax
        self.mfilt_t, self.mfilt_e, self.mfilt_b = mfilt_t, mfilt_e, mfilt_b
        self.lib_dir = ivfs.lib_dir

    def hashdict(self):
        return {'ivfs': self.ivfs.hashdict(), 'filt_t': utils.clhash(self.mfilt_t[:self.lmax + 1]),
                'filt_e': utils.clhash(self.mfilt_e[:self.lmax + 1]), 'filt_b': utils.clhash(self.mfilt_b[:self.lmax + 1])}

    def get_fmask(self):
        return self.ivfs.get_fmask()

    def get_tal(self, a):
        return self.ivfs.get_tal(a)

    @staticmethod
    def almxfm(alm, fm, lmax):
        """alm truncated / padded to lmax with every m column scaled by fm[m]"""
        ret = utils.alm_copy(alm, lmax=lmax)
        start = 0
        for m in range(lmax + 1):           # healpy order: m-major blocks of lmax + 1 - m entries
            ret[start:start + lmax + 1 - m] *= fm[m]
            start += lmax + 1 - m
        return ret

    def _l_rescal(self, fm):
        w = 2 * np.cumsum(fm[:self.lmax + 1]) - fm[0]
        return np.sqrt(w / (2 * np.arange(self.lmax + 1) + 1))

    def get_ftl(self):
        return self.ivfs.get_ftl()[:self.lmax + 1] * self._l_rescal(self.mfilt_t)

    def get_fel(self):
        return self.ivfs.get_fel()[:self.lmax + 1] * self._l_rescal(self.mfilt_e)

    def get_fbl(self):
        return self.ivfs.get_fbl()[:self.lmax + 1] * self._l_rescal(self.mfilt_b)

    def get_sim_tlm(self, idx):
        return self.almxfm(self.ivfs.get_sim_tlm(idx), self.mfilt_t, self.lmax)

    # the reference scales the inverse-variance filtered E and B with the *temperature* weights (filt_util.py:169-173)
    def get_sim_elm(self, idx):
        return self.almxfm(self.ivfs.get_sim_elm(idx), self.mfilt_t, self.lmax)

    def get_sim_blm(self, idx):
        return self.almxfm(self.ivfs.get_sim_blm(idx), self.mfilt_t, self.lmax)

    def get_sim_tmliklm(self, idx):
        return self.almxfm(self.ivfs.get_sim_tmliklm(idx), self.mfilt_t, self.lmax)

    def get_sim_emliklm(self, idx):
        return self.almxfm(self.ivfs.get_sim_emliklm(idx), self.mfilt_e, self.lmax)

    def get_sim_bmliklm(self, idx):
        return self.almxfm(self.ivfs.get_sim_bmliklm(idx), self.mfilt_b, self.lmax)


def _alm_copy(alm, mmaxin, lmaxout, mmaxout):
    """Copy of a healpy alm array with new lmax and mmax (reference: filt_util.py:10-37)."""
    lmaxin = hp.Alm.getlmax(alm.size, mmaxin)
    if mmaxin is None or mmaxin < 0:
        mmaxin = lmaxin
    if lmaxin == lmaxout and mmaxin == mmaxout:
        return np.copy(alm)
    ret = np.zeros(hp.Alm.getsize(lmaxout, mmaxout), dtype=complex)
    n = min(lmaxout, lmaxin) + 1
    for m in range(min(mmaxout, mmaxin) + 1):
        i = m * (2 * lmaxin + 1 - m) // 2 + m
        o = m * (2 * lmaxout + 1 - m) // 2 + m
        ret[o:o + n - m] = alm[i:i + n - m]
    return ret


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

    def __getattr__(self, name):
        if name in ('get_sim_teblm_dev', 'get_sim_mliklm_dev') and hasattr(self.ivfs, name):
            inner = getattr(self.ivfs, name)
            return lambda idx, fields='teb': inner(self.idxs[idx], fields)
        if name == 'flush' and hasattr(self.ivfs, 'flush'):
            return self.ivfs.flush
        if name == 'prefetch_dev' and hasattr(self.ivfs, name):
            return lambda idxs: self.ivfs.prefetch_dev([self.idxs[i] for i in idxs])
        raise AttributeError(name)

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
