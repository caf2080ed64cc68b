"""Gaussian CMB alm simulations from fiducial spectra (reference: plancklens/sims/cmbs.py:25-102)."""
import numpy as np

from .. import hp, utils


def _get_fields(cls):
    order = ['p', 't', 'e', 'b', 'o']
    ret = [f for f in order if (f + f) in cls.keys()]
    for k in cls.keys():
        for f in k:
            if f not in ret:
                ret.append(f)
    return ret


class sims_cmb_unl:
    """Correlated Gaussian alms: phases coloured by the per-l matrix square root of the spectra."""

    def __init__(self, cls_unl, lib_pha):
        lmax = lib_pha.lmax
        fields = _get_fields(cls_unl)
        nf = len(fields)
        rmat = np.zeros((lmax + 1, nf, nf))
        for i, a in enumerate(fields):
            for j, b in enumerate(fields):
                if j >= i and (a + b) in cls_unl:
                    rmat[:, i, j] = cls_unl[a + b][:lmax + 1]
                    rmat[:, j, i] = rmat[:, i, j]
        t, v = np.linalg.eigh(rmat)
        assert np.all(t >= -1e-14 * np.max(np.abs(t))), 'spectra not positive semi-definite'
        t = np.maximum(t, 0.)
        self.rmat = np.einsum('lij,lj,lkj->lik', v, np.sqrt(t), v)
        self._cl_hash = {k: utils.clhash(cls_unl[k]) for k in cls_unl.keys()}
        self.lmax = lmax
        self.lib_pha = lib_pha
        self.fields = fields

    def hashdict(self):
        ret = dict(self._cl_hash)
        ret['phas'] = self.lib_pha.hashdict()
        return ret

    def _phases(self, idx):
        """unit phases of all fields of simulation idx; the last simulation drawn is kept (T, E and B of one sky are
        asked for one after the other, and each needs every phase field)"""
        if getattr(self, '_pha_idx', None) != idx:
            self._pha = [self.lib_pha.get_sim(idx, idf=i) for i in range(len(self.fields))]
            self._pha_idx = idx
        return self._pha

    def _get_sim_alm(self, idx, idf):
        pha = self._phases(idx)
        ret = hp.almxfl(pha[0], self.rmat[:, idf, 0])
        for i in range(1, len(self.fields)):
            if np.any(self.rmat[:, idf, i]):
                ret += hp.almxfl(pha[i], self.rmat[:, idf, i])
        return ret

    def get_sim_alm(self, idx, field):
        assert field in self.fields, self.fields
        return self._get_sim_alm(idx, self.fields.index(field))

    # ---- the same on the device (phase library built with device=True): phases drawn by the Philox kernel, coloured by
    #      one alm_combine launch per field; nothing touches the host
    def _phases_dev(self, idx):
        if getattr(self, '_dpha_idx', None) != idx:
            self._dpha = [self.lib_pha.get_sim_dev(idx, i) for i in range(len(self.fields))]
            self._dpha_idx = idx
        return self._dpha

    def get_sim_alm_dev(self, idx, field):
        """complex128 CUDA tensor of field `field` of simulation idx"""
        from .. import sht
        assert field in self.fields, self.fields
        idf = self.fields.index(field)
        if not hasattr(self, '_rmat_d'):
            self._rmat_d = {}
        terms = []
        for i, ph in enumerate(self._phases_dev(idx)):
            if np.any(self.rmat[:, idf, i]):
                if (idf, i) not in self._rmat_d:
                    self._rmat_d[(idf, i)] = sht.dev_fl(self.rmat[:, idf, i], self.lmax)
                terms.append((ph, self._rmat_d[(idf, i)]))
        assert 1 <= len(terms) <= 4, len(terms)
        return sht.alm_combine(terms)

    def has_device_sims(self):
        return bool(getattr(self.lib_pha, 'device', False))

    def get_sim_tlm(self, idx):
        return self.get_sim_alm(idx, 't')

    def get_sim_elm(self, idx):
        return self.get_sim_alm(idx, 'e')

    def get_sim_blm(self, idx):
        return self.get_sim_alm(idx, 'b')

    def get_sim_plm(self, idx):
        """lensing potential (field 'p' of the input spectra, e.g. from 'pp', 'pt', 'pe')"""
        return self.get_sim_alm(idx, 'p')

    def get_sim_olm(self, idx):
        """lensing curl potential (field 'o')"""
        return self.get_sim_alm(idx, 'o')

    def get_sim_alms(self, idx):
        """all fields of one simulation, in the order of `self.fields` (the reference's version, cmbs.py:95-101,
        builds the same array but forgets to return it)"""
        return np.array([self._get_sim_alm(idx, i) for i in range(len(self.fields))])
