"""CMB + noise map simulations (reference: plancklens/sims/maps.py:10-173); the synthesis runs on the GPU."""
import numpy as np

from .. import hp
from ..utils import clhash
from . import phas


class cmb_maps(object):
    """Lensed-CMB alm library x transfer function -> maps (reference: maps.py:10-91)."""

    def __init__(self, sims_cmb_len, cl_transf, nside=2048, cl_transf_P=None, lib_dir=None):
        self.sims_cmb_len = sims_cmb_len
        self.cl_transf_T = cl_transf
        self.cl_transf_P = np.copy(cl_transf) if cl_transf_P is None else cl_transf_P
        self.nside = nside
        # device_maps = True: get_sim_tmap / get_sim_pmap return float64 CUDA tensors (signal synthesised and noise
        # added on the GPU) which the cinv filters take as they are -- no 400 MB round trip through the host per map
        self.device_maps = False

    def hashdict(self):
        ret = {'sims_cmb_len': self.sims_cmb_len.hashdict(), 'nside': self.nside, 'cl_transf': clhash(self.cl_transf_T)}
        if not np.all(self.cl_transf_P == self.cl_transf_T):
            ret['cl_transf_P'] = clhash(self.cl_transf_P)
        return ret

    def get_sim_tmap(self, idx):
        if self.device_maps:
            return self.get_sim_tmap_dev(idx)
        tlm = hp.almxfl(self.sims_cmb_len.get_sim_tlm(idx), self.cl_transf_T)
        return hp.alm2map(tlm, self.nside) + self.get_sim_tnoise(idx)

    def get_sim_pmap(self, idx):
        if self.device_maps:
            return self.get_sim_pmap_dev(idx)
        elm = hp.almxfl(self.sims_cmb_len.get_sim_elm(idx), self.cl_transf_P)
        blm = hp.almxfl(self.sims_cmb_len.get_sim_blm(idx), self.cl_transf_P)
        Q, U = hp.alm2map_spin([elm, blm], self.nside, 2, hp.Alm.getlmax(elm.size))
        return Q + self.get_sim_qnoise(idx), U + self.get_sim_unoise(idx)

    # ---- device-resident simulated maps (float64 CUDA tensors): CMB alms from the device when the CMB library draws
    #      there, transfer function fused into the synthesis, noise added by the generator kernel itself
    def _cmb_alm_dev(self, idx, field):
        from .. import sht
        lib = self.sims_cmb_len
        if hasattr(lib, 'get_sim_alm_dev') and lib.has_device_sims():
            return lib.get_sim_alm_dev(idx, field)
        return sht.dev_alm(getattr(lib, 'get_sim_%slm' % field)(idx))

    def _transf_dev(self, which, lmax):
        from .. import sht
        if not hasattr(self, '_tf_d'):
            self._tf_d = {}
        if (which, lmax) not in self._tf_d:
            self._tf_d[(which, lmax)] = sht.dev_fl(self.cl_transf_T if which == 't' else self.cl_transf_P, lmax)
        return self._tf_d[(which, lmax)]

    def get_sim_tmap_dev(self, idx):
        from .. import sht
        tlm = self._cmb_alm_dev(idx, 't')
        lmax = sht.alm_lmax(tlm.numel())
        t = sht.get_plan(self.nside, lmax).alm2map(tlm, fl=self._transf_dev('t', lmax))
        return self._add_noise_dev(t, idx, 't')

    def get_sim_pmap_dev(self, idx):
        from .. import sht
        elm, blm = self._cmb_alm_dev(idx, 'e'), self._cmb_alm_dev(idx, 'b')
        lmax = sht.alm_lmax(elm.numel())
        fl = self._transf_dev('p', lmax)
        Q, U = sht.get_plan(self.nside, lmax).alm2map_spin(elm, blm, 2, flg=fl, flc=fl)
        return self._add_noise_dev(Q, idx, 'q'), self._add_noise_dev(U, idx, 'u')

    def _add_noise_dev(self, m, idx, field):
        """m += noise map of `field` in ('t', 'q', 'u') on the device"""
        from .. import sht
        noise = {'t': self.get_sim_tnoise, 'q': self.get_sim_qnoise, 'u': self.get_sim_unoise}[field](idx)
        return sht.map_axpy(m, sht.dev_map(noise), 1.0)

    def get_sim_tnoise(self, idx):
        assert 0, 'subclass this'

    def get_sim_qnoise(self, idx):
        assert 0, 'subclass this'

    def get_sim_unoise(self, idx):
        assert 0, 'subclass this'


class cmb_maps_noisefree(cmb_maps):
    def get_sim_tnoise(self, idx):
        return np.zeros(hp.nside2npix(self.nside))

    get_sim_qnoise = get_sim_tnoise
    get_sim_unoise = get_sim_tnoise


class cmb_maps_nlev(cmb_maps):
    """Homogeneous white noise of nlev_t / nlev_p uK-arcmin on top (reference: maps.py:116-173)."""

    def __init__(self, sims_cmb_len, cl_transf, nlev_t, nlev_p, nside, lib_dir=None, pix_lib_phas=None):
        if pix_lib_phas is None:
            assert lib_dir is not None
            pix_lib_phas = phas.pix_lib_phas(lib_dir, 3, (hp.nside2npix(nside),))
        assert pix_lib_phas.shape == (hp.nside2npix(nside),), (pix_lib_phas.shape, (hp.nside2npix(nside),))
        self.pix_lib_phas = pix_lib_phas
        self.nlev_t = nlev_t
        self.nlev_p = nlev_p
        super(cmb_maps_nlev, self).__init__(sims_cmb_len, cl_transf, nside=nside, lib_dir=lib_dir)

    def hashdict(self):
        ret = super(cmb_maps_nlev, self).hashdict()
        ret.update({'nlev_t': self.nlev_t, 'nlev_p': self.nlev_p, 'pixphas': self.pix_lib_phas.hashdict()})
        return ret

    def _vamin(self):
        return np.sqrt(hp.nside2pixarea(self.nside, degrees=True)) * 60

    def _add_noise_dev(self, m, idx, field):
        from .. import sht
        idf, nlev = {'t': (0, self.nlev_t), 'q': (1, self.nlev_p), 'u': (2, self.nlev_p)}[field]
        if getattr(self.pix_lib_phas, 'device', False):
            # drawn on the device: map + nlev / vamin * phase in the generator kernel's own pass over the pixels
            return self.pix_lib_phas.get_sim_dev(idx, idf, scale=nlev / self._vamin(), add=m).reshape(-1)
        # host phases go to the device as drawn; the nlev / vamin scaling is the axpy coefficient
        return sht.map_axpy(m, sht.dev_map(self.pix_lib_phas.get_sim(idx, idf=idf)), nlev / self._vamin())

    def get_sim_tnoise(self, idx):
        return self.nlev_t / self._vamin() * self.pix_lib_phas.get_sim(idx, idf=0)

    def get_sim_qnoise(self, idx):
        return self.nlev_p / self._vamin() * self.pix_lib_phas.get_sim(idx, idf=1)

    def get_sim_unoise(self, idx):
        return self.nlev_p / self._vamin() * self.pix_lib_phas.get_sim(idx, idf=2)


class cmb_maps_harmonicspace(object):
    r"""Simulations made directly in harmonic space: transfer function times the lensed CMB alms plus statistically
    isotropic (possibly non-white) noise drawn from a harmonic phase library (reference: maps.py:176-275).

        Args:
            sims_cmb_len: lensed CMB library
            cls_transf: transfer functions for 't', 'e', 'b'
            cls_noise: noise spectra for 't', 'e', 'b'
            noise_phas: `phas.lib_phas` with at least three fields, same lmax as the CMB library
            lib_dir (optional): the hash is checked against a cached copy there
            nside (optional): maps are returned in pixel space at this resolution instead of as alms
    """

    def __init__(self, sims_cmb_len, cls_transf, cls_noise, noise_phas, lib_dir=None, nside=None):
        import os
        import pickle as pk
        from ..helpers import mpi
        from ..utils import hash_check
        assert noise_phas.nfields >= 3, noise_phas.nfields
        self.sims_cmb_len, self.cls_transf, self.cls_noise = sims_cmb_len, cls_transf, cls_noise
        self.phas, self.nside = noise_phas, nside
        if hasattr(sims_cmb_len, 'lmax'):
            assert sims_cmb_len.lmax == noise_phas.lmax, (sims_cmb_len.lmax, noise_phas.lmax)
        if lib_dir is not None:
            fn_hash = os.path.join(lib_dir, 'sim_hash.pk')
            if mpi.rank == 0 and not os.path.exists(fn_hash):
                os.makedirs(lib_dir, exist_ok=True)
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
            mpi.barrier()
            with open(fn_hash, 'rb') as f:
                hash_check(self.hashdict(), pk.load(f))

    def hashdict(self):
        ret = {'sims_cmb_len': self.sims_cmb_len.hashdict(), 'phas': self.phas.hashdict()}
        ret.update({'noise' + k: clhash(v) for k, v in self.cls_noise.items()})
        ret.update({'transf' + k: clhash(v) for k, v in self.cls_transf.items()})
        return ret

    def _noise(self, idx, a):
        assert a in self.cls_noise, a
        return hp.almxfl(self.phas.get_sim(idx, 'teb'.index(a)), np.sqrt(self.cls_noise[a]))

    def get_sim_tnoise(self, idx):
        return self._noise(idx, 't')

    def get_sim_enoise(self, idx):
        return self._noise(idx, 'e')

    def get_sim_bnoise(self, idx):
        return self._noise(idx, 'b')

    def get_sim_tmap(self, idx):
        """temperature alms, or the map if nside was given"""
        assert 't' in self.cls_transf
        tlm = hp.almxfl(self.sims_cmb_len.get_sim_tlm(idx), self.cls_transf['t']) + self.get_sim_tnoise(idx)
        return hp.alm2map(tlm, self.nside) if self.nside else tlm

    def get_sim_pmap(self, idx):
        """(elm, blm), or the (Q, U) maps if nside was given"""
        assert 'e' in self.cls_transf and 'b' in self.cls_transf
        elm = hp.almxfl(self.sims_cmb_len.get_sim_elm(idx), self.cls_transf['e']) + self.get_sim_enoise(idx)
        blm = hp.almxfl(self.sims_cmb_len.get_sim_blm(idx), self.cls_transf['b']) + self.get_sim_bnoise(idx)
        if self.nside is not None:
            return hp.alm2map_spin([elm, blm], self.nside, 2, hp.Alm.getlmax(elm.size))
        return elm, blm
