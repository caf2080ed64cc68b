"""Conjugate-gradient inverse-variance filtering libraries (reference: plancklens/filt/filt_cinv.py).

`cinv_t` / `cinv_p` own a multigrid-preconditioned CG chain whose operators run on the GPU
(`plancklens_b200.qcinv`); `library_cinv_sepTP` plugs them into the cached-library interface of
`filt_simple.library_sepTP`.  Default chains are the reference's (filt_cinv.py:113-116, :237-239).
"""
import os
import pickle as pk

import numpy as np

from .. import hp, utils
from ..helpers import mpi
from ..qcinv import cd_solve, multigrid, opfilt_pp, opfilt_tp, opfilt_tt, util, util_alm
from . import filt_simple


class cinv(object):
    def __init__(self, lib_dir, lmax):
        self.lib_dir = lib_dir
        self.lmax = lmax

    def _load(self, name, lmax):
        lmax = self.lmax if lmax is None else lmax
        ret = np.loadtxt(os.path.join(self.lib_dir, name))
        assert len(ret) > lmax, (len(ret), lmax)
        return ret[:lmax + 1]

    def get_tal(self, a, lmax=None):
        assert a.lower() in ['t', 'e', 'b'], a
        return self._load("tal.dat", lmax)

    def get_fmask(self):
        return hp.read_map(os.path.join(self.lib_dir, "fmask.fits.gz"))

    def get_ftl(self, lmax=None):
        return self._load("ftl.dat", lmax)

    def get_fel(self, lmax=None):
        return self._load("fel.dat", lmax)

    def get_fbl(self, lmax=None):
        return self._load("fbl.dat", lmax)


def _ninv_hash(comps):
    return [utils.clhash(c) if isinstance(c, np.ndarray) and c.size > 1 else c for c in comps]


class cinv_t(cinv):
    r"""Temperature-only inverse-variance (Wiener) filter (reference: filt_cinv.py:56-203).

        Args:
            lib_dir: mask and isotropic approximations are cached there
            lmax: filtered alm's are reconstructed up to lmax
            nside: resolution of the maps to filter
            cl: fiducial CMB spectra (dict with 'tt')
            transf: transfer function
            ninv: list of maps / paths whose product is the inverse pixel variance
            rescal_cl: isotropic rescaling of the unknowns before the CG (default: :math:`\sqrt{\ell(\ell+1)/2\pi}`),
                       which changes the convergence criterion only
    """

    def __init__(self, lib_dir, lmax, nside, cl, transf, ninv, rescal_cl='default', marge_monopole=True,
                 marge_dipole=True, marge_maps=(), pcf='default', chain_descr=None):
        assert lib_dir is not None and lmax >= 1024 and nside >= 512, (lib_dir, lmax, nside)
        assert isinstance(ninv, list)
        super(cinv_t, self).__init__(lib_dir, lmax)
        if rescal_cl in ['default', None]:
            default_rescal = True
            rescal_cl = np.sqrt(np.arange(lmax + 1, dtype=float) * np.arange(1, lmax + 2, dtype=float) / 2. / np.pi)
        else:
            default_rescal = False
            assert len(rescal_cl) >= lmax + 1, [rescal_cl.shape, lmax]
        dl = {k: rescal_cl[:lmax + 1] ** 2 * cl[k][:lmax + 1] for k in cl.keys()}
        transf_dl = transf[:lmax + 1] * utils.cli(rescal_cl)
        self.nside = nside
        self.cl = cl
        self.dl = dl
        self.transf = transf[:lmax + 1]
        self.rescaled_transf = transf_dl
        self.rescal_cl = rescal_cl
        self.default_rescal = default_rescal
        self.ninv = ninv
        self.marge_monopole = marge_monopole
        self.marge_dipole = marge_dipole
        self.marge_maps = marge_maps

        pcf = os.path.join(lib_dir, "dense.pk") if pcf == 'default' else ''
        if chain_descr is None:
            chain_descr = \
                [[3, ["split(dense(" + pcf + "), 64, diag_cl)"], 256, 128, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [2, ["split(stage(3),  256, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [1, ["split(stage(2),  512, diag_cl)"], 1024, 512, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [0, ["split(stage(1), 1024, diag_cl)"], lmax, nside, np.inf, 1.0e-5, cd_solve.tr_cg, cd_solve.cache_mem()]]
        n_inv_filt = util.jit(opfilt_tt.alm_filter_ninv, ninv, transf_dl, marge_monopole=marge_monopole,
                              marge_dipole=marge_dipole, marge_maps=marge_maps)
        self.chain_descr = chain_descr
        self.chain = util.jit(multigrid.multigrid_chain, opfilt_tt, self.chain_descr, dl, n_inv_filt)
        # B200: this filter's solves run in lane 2, the polarization filter's in lane 1 (own transform plans and reduction
        # scratch, `sht.use_lane`); lane 0 stays with the caller (simulated maps, quadratic estimators), so that
        # library_cinv_sepTP can run all three at once
        self.lane = 2
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            fn = os.path.join(lib_dir, "filt_hash.pk")
            if not os.path.exists(fn):
                with open(fn, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
            if not os.path.exists(os.path.join(lib_dir, "ftl.dat")):
                np.savetxt(os.path.join(lib_dir, "ftl.dat"), self._calc_ftl())
            if not os.path.exists(os.path.join(lib_dir, "tal.dat")):
                np.savetxt(os.path.join(lib_dir, "tal.dat"), self._calc_tal())
            if not os.path.exists(os.path.join(lib_dir, "fmask.fits.gz")):
                hp.write_map(os.path.join(lib_dir, "fmask.fits.gz"), self._calc_mask())
        mpi.barrier()
        fn = os.path.join(lib_dir, "filt_hash.pk")
        with open(fn, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=fn)

    def _calc_ftl(self):
        ninv = self.chain.n_inv_filt.n_inv
        npix = len(ninv)
        NlevT_uKamin = np.sqrt(4. * np.pi / npix / np.sum(ninv) * len(np.where(ninv != 0.0)[0])) * 180. * 60. / np.pi
        print("cinv_t::noiseT_uk_arcmin = %.3f" % NlevT_uKamin)
        s_cls = self.cl
        if s_cls['tt'][0] == 0.:
            assert self.chain.n_inv_filt.marge_monopole
        if s_cls['tt'][1] == 0.:
            assert self.chain.n_inv_filt.marge_dipole
        ftl = utils.cli(s_cls['tt'][0:self.lmax + 1]
                        + (NlevT_uKamin * np.pi / 180. / 60.) ** 2 * utils.cli(self.transf[0:self.lmax + 1] ** 2))
        if self.chain.n_inv_filt.marge_monopole:
            ftl[0] = 0.0
        if self.chain.n_inv_filt.marge_dipole:
            ftl[1] = 0.0
        return ftl

    def _calc_tal(self):
        return utils.cli(self.transf)

    def _calc_mask(self):
        ninv = self.chain.n_inv_filt.n_inv
        assert hp.npix2nside(len(ninv)) == self.nside
        return np.where(ninv > 0, 1., 0.)

    def hashdict(self):
        hd = {'lmax': self.lmax, 'nside': self.nside, 'cltt': utils.clhash(self.cl['tt'][:self.lmax + 1]),
              'transf': utils.clhash(self.transf[:self.lmax + 1]), 'ninv': _ninv_hash(self.ninv),
              'marge_monopole': self.marge_monopole, 'marge_dipole': self.marge_dipole,
              'marge_maps': self.marge_maps}
        if self.default_rescal is False:
            hd['rescal_cl'] = utils.clhash(self.rescal_cl)
        return hd

    def apply_ivf(self, tmap, soltn=None):
        """Inverse-variance filtered alm of a temperature map (numpy in, numpy out; the solve runs on the GPU)."""
        return self.apply_ivf_dev(tmap, soltn=soltn).cpu().numpy()

    def apply_ivf_dev(self, tmap, soltn=None):
        """Same with the result left on the device (complex128 CUDA tensor); `tmap` and `soltn` may be numpy arrays or
        CUDA tensors.  What `library_cinv_sepTP.get_sim_teblm_dev` hands to `qest` without a host round trip."""
        from .. import sht
        if soltn is None:
            talm = util_alm.dalm.zeros(self.lmax)
        else:
            talm = util_alm.dalm(sht.dev_alm(soltn).clone())
        with sht.use_lane(self.lane):
            self.chain.solve(talm, tmap)
        if not hasattr(self, '_rescal_d'):
            self._rescal_d = sht.dev_fl(self.rescal_cl, self.lmax)
        return talm.almxfl(self._rescal_d, inplace=True).t


class cinv_p(cinv):
    r"""Polarization-only inverse-variance (Wiener) filter (reference: filt_cinv.py:206-338).

        ninv: list of 1 (QQ = UU) or 3 (QQ, QU, UU) lists of maps / paths.
    """

    def __init__(self, lib_dir, lmax, nside, cl, transf, ninv, pcf='default', chain_descr=None, transf_blm=None,
                 marge_qmaps=(), marge_umaps=()):
        assert lib_dir is not None and lmax >= 1024 and nside >= 512, (lib_dir, lmax, nside)
        super(cinv_p, self).__init__(lib_dir, lmax)
        self.nside = nside
        self.cl = cl
        self.transf_e = transf
        self.transf_b = transf if transf_blm is None else transf_blm
        self.transf = transf if transf_blm is None else 0.5 * self.transf_e + 0.5 * self.transf_b
        self.ninv = ninv
        pcf = os.path.join(lib_dir, "dense.pk") if pcf == 'default' else ''
        if chain_descr is None:
            chain_descr = \
                [[2, ["split(dense(" + pcf + "), 32, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [1, ["split(stage(2),  512, diag_cl)"], 1024, 512, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [0, ["split(stage(1), 1024, diag_cl)"], lmax, nside, np.inf, 1.0e-5, cd_solve.tr_cg, cd_solve.cache_mem()]]
        n_inv_filt = util.jit(opfilt_pp.alm_filter_ninv, ninv, transf[0:lmax + 1], b_transf_b=transf_blm,
                              marge_umaps=marge_umaps, marge_qmaps=marge_qmaps)
        self.chain_descr = chain_descr
        self.chain = util.jit(multigrid.multigrid_chain, opfilt_pp, chain_descr, cl, n_inv_filt)
        # B200: this filter's solves run in lane 1 (see cinv_t)
        self.lane = 1
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            fn = os.path.join(lib_dir, "filt_hash.pk")
            if not os.path.exists(fn):
                with open(fn, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
            if not os.path.exists(os.path.join(lib_dir, "fbl.dat")):
                fel, fbl = self._calc_febl()
                np.savetxt(os.path.join(lib_dir, "fel.dat"), fel)
                np.savetxt(os.path.join(lib_dir, "fbl.dat"), fbl)
            if not os.path.exists(os.path.join(lib_dir, "tal.dat")):
                np.savetxt(os.path.join(lib_dir, "tal.dat"), self._calc_tal())
            if not os.path.exists(os.path.join(lib_dir, "fmask.fits.gz")):
                hp.write_map(os.path.join(lib_dir, "fmask.fits.gz"), self._calc_mask())
        mpi.barrier()
        fn = os.path.join(lib_dir, "filt_hash.pk")
        with open(fn, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=fn)

    def hashdict(self):
        return {'lmax': self.lmax, 'nside': self.nside,
                'clee': utils.clhash(self.cl.get('ee', np.array([0.]))),
                'cleb': utils.clhash(self.cl.get('eb', np.array([0.]))),
                'clbb': utils.clhash(self.cl.get('bb', np.array([0.]))),
                'transf': utils.clhash(self.transf), 'ninv': [_ninv_hash(self.ninv[0])]}

    def apply_ivf(self, tmap, soltn=None):
        """Inverse-variance filtered (E, B) alms of a (Q, U) map pair."""
        e, b = self.apply_ivf_dev(tmap, soltn=soltn)
        return e.cpu().numpy(), b.cpu().numpy()

    def apply_ivf_dev(self, tmap, soltn=None):
        """Same with the (E, B) results left on the device; maps and `soltn` may be numpy arrays or CUDA tensors."""
        from .. import sht
        if soltn is not None:
            assert len(soltn) == 2
            assert sht.alm_lmax(len(soltn[0])) == self.lmax and sht.alm_lmax(len(soltn[1])) == self.lmax
            talm = util_alm.eblm([util_alm.dalm(sht.dev_alm(soltn[0]).clone()), util_alm.dalm(sht.dev_alm(soltn[1]).clone())])
        else:
            talm = util_alm.eblm([util_alm.dalm.zeros(self.lmax), util_alm.dalm.zeros(self.lmax)])
        assert len(tmap) == 2
        with sht.use_lane(self.lane):
            self.chain.solve(talm, [tmap[0], tmap[1]])
        return talm.elm.t, talm.blm.t

    def _calc_febl(self):
        assert 'eb' not in self.chain.s_cls.keys()
        ninv = self.chain.n_inv_filt.get_ninv()
        lev = lambda n: np.sqrt(4. * np.pi / len(n) / np.sum(n) * len(np.where(n != 0.0)[0])) * 180. * 60. / np.pi
        if len(ninv) == 1:
            NlevP_uKamin = lev(ninv[0])
        else:
            assert len(ninv) == 3
            NlevP_uKamin = 0.5 * lev(ninv[0]) + 0.5 * lev(ninv[2])
        print("cinv_p::noiseP_uk_arcmin = %.3f" % NlevP_uKamin)
        s_cls = self.chain.s_cls
        b_e = self.chain.n_inv_filt.b_transf_e
        b_b = self.chain.n_inv_filt.b_transf_b
        fel = utils.cli(s_cls['ee'][:self.lmax + 1] + (NlevP_uKamin * np.pi / 180. / 60.) ** 2 * utils.cli(b_e[0:self.lmax + 1] ** 2))
        fbl = utils.cli(s_cls['bb'][:self.lmax + 1] + (NlevP_uKamin * np.pi / 180. / 60.) ** 2 * utils.cli(b_b[0:self.lmax + 1] ** 2))
        fel[0:2] *= 0.0
        fbl[0:2] *= 0.0
        return fel, fbl

    def _calc_tal(self):
        return utils.cli(self.transf)

    def _calc_mask(self):
        mask = np.ones(hp.nside2npix(self.nside), dtype=float)
        for ninv in self.chain.n_inv_filt.get_ninv():
            assert hp.npix2nside(len(ninv)) == self.nside
            mask *= (ninv > 0.)
        return mask


class cinv_tp:
    """Joint temperature + polarization inverse-variance (Wiener) filter (reference: filt_cinv.py:341-512).

    Args as the reference: `ninv` = [TT, (QQ+UU)/2] or [TT, QQ, QU, UU] (each a list of maps / numbers to multiply),
    `marge_maps_t`, `marge_monopole`, `marge_dipole` for the temperature block, `rescal_cl` in
    ('default', None, 'tonly'), `transf_p` if the polarization beam differs."""

    def __init__(self, lib_dir, lmax, nside, cl, transf, ninv, marge_maps_t=(), marge_monopole=False, marge_dipole=False,
                 pcf='default', rescal_cl='default', chain_descr=None, transf_p=None):
        assert lmax >= 1024 and nside >= 512, (lmax, nside)
        assert len(ninv) == 2 or len(ninv) == 4
        ls = np.arange(lmax + 1, dtype=float)
        dl_fac = np.sqrt(ls * (ls + 1.) / 2. / np.pi)
        if rescal_cl == 'default':
            rescal_cl = {a: dl_fac.copy() for a in ['t', 'e', 'b']}
        elif rescal_cl is None:
            rescal_cl = {a: np.ones(lmax + 1, dtype=float) for a in ['t', 'e', 'b']}
        elif rescal_cl == 'tonly':
            rescal_cl = {a: np.ones(lmax + 1, dtype=float) for a in ['e', 'b']}
            rescal_cl['t'] = dl_fac.copy()
        else:
            assert 0
        for k in rescal_cl.keys():
            rescal_cl[k] /= np.mean(rescal_cl[k])      # keeps the relative TEB weights of the spectra
        dl = {k: rescal_cl[k[0]] * rescal_cl[k[1]] * cl[k][:lmax + 1] for k in cl.keys()}
        if transf_p is None:
            transf_p = transf
        transf_dls = {a: transf_p[:lmax + 1] * utils.cli(rescal_cl[a]) for a in ['e', 'b']}
        transf_dls['t'] = transf[:lmax + 1] * utils.cli(rescal_cl['t'])
        self.lmax, self.nside, self.cl = lmax, nside, cl
        self.transf_t, self.transf_p = transf, transf_p
        self.ninv = ninv
        self.marge_maps_t = marge_maps_t
        self.marge_maps_p = []
        self.lib_dir = lib_dir
        self.rescal_cl = rescal_cl
        if chain_descr is None:
            pcf = os.path.join(lib_dir, "dense_tp.pk") if pcf == 'default' else ''
            chain_descr = \
                [[3, ["split(dense(" + pcf + "), 64, diag_cl)"], 256, 128, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [2, ["split(stage(3),  256, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [1, ["split(stage(2),  512, diag_cl)"], 1024, 512, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [0, ["split(stage(1), 1024, diag_cl)"], lmax, nside, np.inf, 1.0e-5, cd_solve.tr_cg, cd_solve.cache_mem()]]
        n_inv_filt = util.jit(opfilt_tp.alm_filter_ninv, ninv, transf_dls['t'], b_transf_e=transf_dls['e'],
                              b_transf_b=transf_dls['b'], marge_maps_t=marge_maps_t, marge_monopole=marge_monopole,
                              marge_dipole=marge_dipole)
        self.chain_descr = chain_descr
        self.chain = util.jit(multigrid.multigrid_chain, opfilt_tp, chain_descr, dl, n_inv_filt)
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            fn = os.path.join(lib_dir, "filt_hash.pk")
            if not os.path.exists(fn):
                with open(fn, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
            fn = os.path.join(lib_dir, "fal.pk")
            if not os.path.exists(fn):
                with open(fn, 'wb') as f:
                    pk.dump(self._calc_fal(), f, protocol=2)
            if not os.path.exists(os.path.join(lib_dir, "fmask.fits.gz")):
                hp.write_map(os.path.join(lib_dir, "fmask.fits.gz"), self.calc_mask())
        mpi.barrier()
        with open(os.path.join(lib_dir, "filt_hash.pk"), 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=os.path.join(lib_dir, "filt_hash.pk"))

    def hashdict(self):
        ret = {'lmax': self.lmax, 'nside': self.nside,
               'rescal_cl': {k: utils.clhash(self.rescal_cl[k]) for k in self.rescal_cl.keys()},
               'cls': {k: utils.clhash(self.cl[k]) for k in self.cl.keys()},
               'transf': utils.clhash(self.transf_t), 'ninv': self._ninv_hash(),
               'marge_maps_t': self.marge_maps_t, 'marge_maps_p': self.marge_maps_p}
        if self.transf_p is not self.transf_t:
            ret['transf_p'] = utils.clhash(self.transf_p)
        return ret

    def get_fal(self, lmax=None):
        with open(os.path.join(self.lib_dir, "fal.pk"), 'rb') as f:
            fal = pk.load(f)
        return fal if lmax is None else {k: v[:lmax + 1] for k, v in fal.items()}

    def _calc_fal(self):
        """Isotropic approximation to the filtering matrix (reference: filt_cinv.py:450-476)."""
        ninv = self.chain.n_inv_filt.n_inv
        assert len(ninv) == 2, 'implement this, easy'
        npix = 12 * self.nside ** 2
        assert ninv[0].size == npix and ninv[1].size == npix
        lev = lambda m: np.sqrt(4. * np.pi / npix / np.sum(m) * len(np.where(m != 0.0)[0])) * 180. * 60. / np.pi
        nlevt, nlevp = lev(ninv[0]), lev(ninv[1])
        print("cinv_tp::noiseT_uk_arcmin = %.3f" % nlevt)
        print("cinv_tp::noiseP_uk_arcmin = %.3f" % nlevp)
        fals = np.zeros((self.lmax + 1, 3, 3), dtype=float)
        for i, a in enumerate(['t', 'e', 'b']):
            for j, b in enumerate(['t', 'e', 'b']):
                fals[:, i, j] = self.cl.get(a + b, self.cl.get(b + a, np.zeros(self.lmax + 1)))[:self.lmax + 1]
        fals[1:, 0, 0] += ((nlevt / 180 / 60 * np.pi) / self.transf_t[1:self.lmax + 1]) ** 2
        fals[2:, 1, 1] += ((nlevp / 180 / 60 * np.pi) / self.transf_p[2:self.lmax + 1]) ** 2
        fals[2:, 2, 2] += ((nlevp / 180 / 60 * np.pi) / self.transf_p[2:self.lmax + 1]) ** 2
        fals = np.linalg.pinv(fals)
        fals_dict = {}
        for i, a in enumerate(['t', 'e', 'b']):
            for j, b in enumerate(['t', 'e', 'b'][i:]):
                if np.any(fals[:, i, i + j]):
                    fals_dict[a + b] = fals[:, i, i + j]
        return fals_dict

    def calc_mask(self):
        mask = np.ones(hp.nside2npix(self.nside), dtype=float)
        for ninv in self.chain.n_inv_filt.n_inv:
            assert hp.npix2nside(len(ninv)) == self.nside
            mask *= (ninv > 0.)
        return mask

    def get_fmask(self):
        return hp.read_map(os.path.join(self.lib_dir, "fmask.fits.gz"))

    def apply_ivf(self, tqumap, soltn=None, apply_fini=''):
        """Inverse-variance filtered (T, E, B) alms of a (T, Q, U) map triple; the solve runs on the GPU."""
        assert len(tqumap) == 3
        if soltn is None:
            talm = util_alm.teblm([util_alm.dalm.zeros(self.lmax) for _ in range(3)])
        else:
            talm = util_alm.teblm([util_alm.dalm.from_numpy(hp.almxfl(s, self.rescal_cl[a])) for s, a in zip(soltn, 'teb')])
        self.chain.solve(talm, [tqumap[0], tqumap[1], tqumap[2]], apply_fini=apply_fini)
        t, e, b = talm.numpy()
        return hp.almxfl(t, self.rescal_cl['t']), hp.almxfl(e, self.rescal_cl['e']), hp.almxfl(b, self.rescal_cl['b'])

    def _ninv_hash(self):
        # lists of maps / numbers per component are hashed element-wise (the reference hashes bare arrays only,
        # filt_cinv.py:503-510, and would compare arrays by value inside lists)
        return [[_ninv_hash(c) if isinstance(c, (list, tuple)) else _ninv_hash([c])[0] for c in self.ninv]]


class library_cinv_jTP(filt_simple.library_jTP):
    """CG inverse-variance filtering of a simulation library, T and P filtered jointly
    (reference: filt_cinv.py:584-627)."""

    def __init__(self, lib_dir, sim_lib, cinv_jtp, cl_weights, soltn_lib=None):
        self.cinv_tp = cinv_jtp
        super(library_cinv_jTP, self).__init__(lib_dir, sim_lib, cl_weights, soltn_lib=soltn_lib)
        if mpi.rank == 0:
            fname_mask = os.path.join(self.lib_dir, "fmask.fits.gz")
            if not os.path.exists(fname_mask):
                hp.write_map(fname_mask, self.cinv_tp.get_fmask())
        mpi.barrier()

    def hashdict(self):
        return {'cinv_tp': self.cinv_tp.hashdict(), 'clw': {k: utils.clhash(self.cl[k]) for k in self.cl.keys()},
                'sim_lib': self.sim_lib.hashdict()}

    def get_fmask(self):
        return hp.read_map(os.path.join(self.lib_dir, "fmask.fits.gz"))

    def get_fal(self, lmax=None):
        return self.cinv_tp.get_fal(lmax=lmax)

    def _apply_ivf(self, tqumap, soltn=None):
        return self.cinv_tp.apply_ivf(tqumap, soltn=soltn)


class library_cinv_sepTP(filt_simple.library_sepTP):
    """CG inverse-variance filtering of a simulation library, T and P filtered separately
    (reference: filt_cinv.py:515-587)."""

    def __init__(self, lib_dir, sim_lib, cinvt, cinvp, cl_weights, soltn_lib=None):
        self.cinv_t = cinvt
        self.cinv_p = cinvp
        self.cg_iterations = {}        # idx -> {'T': n, 'P': n}: top-level CG iterations of the solves run by this library
        super(library_cinv_sepTP, self).__init__(lib_dir, sim_lib, cl_weights, soltn_lib=soltn_lib)
        if mpi.rank == 0:
            fname_mask = os.path.join(self.lib_dir, "fmask.fits.gz")
            if not os.path.exists(fname_mask):
                fmask = self.cinv_t.get_fmask()
                assert np.all(fmask == self.cinv_p.get_fmask())
                hp.write_map(fname_mask, fmask)
        mpi.barrier()

    def hashdict(self):
        return {'cinv_t': self.cinv_t.hashdict(), 'cinv_p': self.cinv_p.hashdict(), 'sim_lib': self.sim_lib.hashdict()}

    def get_fmask(self):
        return hp.read_map(os.path.join(self.lib_dir, "fmask.fits.gz"))

    def get_tal(self, a, lmax=None):
        assert a.lower() in ['t', 'e', 'b'], a
        return (self.cinv_t if a.lower() == 't' else self.cinv_p).get_tal(a, lmax=lmax)

    def get_ftl(self, lmax=None):
        return self.cinv_t.get_ftl(lmax=lmax)

    def get_fel(self, lmax=None):
        return self.cinv_p.get_fel(lmax=lmax)

    def get_fbl(self, lmax=None):
        return self.cinv_p.get_fbl(lmax=lmax)

    def _apply_ivf_t(self, tmap, soltn=None):
        return self.cinv_t.apply_ivf(tmap, soltn=soltn)

    def _apply_ivf_p(self, pmap, soltn=None):
        return self.cinv_p.apply_ivf(pmap, soltn=soltn)

    _TP_CONCURRENT = True          # the T and P solves of one simulation run side by side (filt_simple.library_sepTP)

    def _tp_ready(self):
        """Side by side only once both chains have captured their preconditioner graphs: the first solves create plans and
        tables (cudaMalloc, synchronous uploads), which CUDA refuses while a stream of the process is capturing."""
        def warm(c):
            return all(not isinstance(op, multigrid.graphed_op) or op.graph is not None for op in c.chain.bstage.pre_ops)
        return warm(self.cinv_t) and warm(self.cinv_p)

    def _filter_t_dev(self, idx, tmap=None):
        out = super(library_cinv_sepTP, self)._filter_t_dev(idx, tmap)
        self.cg_iterations.setdefault(idx, {})['T'] = int(self.cinv_t.chain.niter)     # read on the lane that solved
        return out

    def _filter_p_dev(self, idx, pmap=None):
        out = super(library_cinv_sepTP, self)._filter_p_dev(idx, pmap)
        self.cg_iterations.setdefault(idx, {})['P'] = int(self.cinv_p.chain.niter)
        return out

    def _apply_ivf_t_dev(self, tmap, soltn=None):
        return self.cinv_t.apply_ivf_dev(tmap, soltn=soltn)

    def _apply_ivf_p_dev(self, pmap, soltn=None):
        return self.cinv_p.apply_ivf_dev(pmap, soltn=soltn)

    def get_tmliklm(self, idx):
        return hp.almxfl(self.get_sim_tlm(idx), self.cinv_t.cl['tt'])

    def get_emliklm(self, idx):
        return hp.almxfl(self.get_sim_elm(idx), self.cinv_p.cl['ee'])

    def get_bmliklm(self, idx):
        return hp.almxfl(self.get_sim_blm(idx), self.cinv_p.cl['bb'])
