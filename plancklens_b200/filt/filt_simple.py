"""Non-iterative CMB filtering libraries (reference: plancklens/filt/filt_simple.py).

`library_sepTP` is the template class every separately-filtered (T and P) library derives from: it caches the
inverse-variance filtered alms on disk and derives the Wiener-filtered ones.  `library_fullsky_sepTP` is the
isotropic full-sky filter of params/idealized_example.py; its transforms run on the GPU.
"""
import os
import pickle as pk

import numpy as np

from .. import hp, utils
from ..helpers import mpi


class library_sepTP(object):
    """Inverse-variance and Wiener filtering of a simulation library, T and P filtered independently
    (reference: filt_simple.py:16-183).

    Args:
        lib_dir: hashes and filtered alms are cached there
        sim_lib: simulation library with `get_sim_tmap`, `get_sim_pmap`
        cl_weights: CMB spectra turning inverse-variance filtered alms into Wiener-filtered ones
    """

    def __init__(self, lib_dir, sim_lib, cl_weights, soltn_lib=None, cache=True):
        self.lib_dir = lib_dir
        self.sim_lib = sim_lib
        self.cl = cl_weights
        self.soltn_lib = soltn_lib
        self.cache = cache
        fn_hash = os.path.join(lib_dir, 'filt_hash.pk')
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            if not os.path.exists(fn_hash):
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fn_hash, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=fn_hash)

    # -- to be provided by the concrete filter
    def hashdict(self):
        assert 0, 'override this'

    def get_fmask(self):
        assert 0, 'override this'

    def _apply_ivf_t(self, tmap, soltn=None):
        assert 0, 'override this'

    def _apply_ivf_p(self, pmap, soltn=None):
        assert 0, 'override this'

    def get_ftl(self):
        r"""Isotropic approximation :math:`F^{T}_\ell = (C_\ell^{TT} + N^{T}_\ell / b_\ell^2)^{-1}`."""
        assert 0, 'override this'

    def get_fel(self):
        assert 0, 'override this'

    def get_fbl(self):
        assert 0, 'override this'

    def get_tal(self, a):
        assert 0, 'override this'

    # -- cached products
    def _fname(self, idx, a):
        return os.path.join(self.lib_dir, ('sim_%04d_%slm.fits' % (idx, a)) if idx >= 0 else 'dat_%slm.fits' % a)

    def get_sim_tlm(self, idx):
        """Inverse-variance filtered temperature alm of simulation idx (idx = -1: data)."""
        fn = self._fname(idx, 't')
        if os.path.exists(fn):
            return hp.read_alm(fn)
        soltn = None if self.soltn_lib is None else self.soltn_lib.get_sim_tmliklm(idx)
        tlm = self._apply_ivf_t(self.sim_lib.get_sim_tmap(idx), soltn=soltn)
        if self.cache:
            hp.write_alm(fn, tlm, overwrite=True)
        return tlm

    def _get_sim_eblm(self, idx, which):
        fn_e, fn_b = self._fname(idx, 'e'), self._fname(idx, 'b')
        fn = fn_e if which == 'e' else fn_b
        if os.path.exists(fn):
            return hp.read_alm(fn)
        soltn = None
        if self.soltn_lib is not None:
            soltn = np.array([self.soltn_lib.get_sim_emliklm(idx), self.soltn_lib.get_sim_bmliklm(idx)])
        elm, blm = self._apply_ivf_p(self.sim_lib.get_sim_pmap(idx), soltn=soltn)
        if self.cache:
            hp.write_alm(fn_e, elm, overwrite=True)
            hp.write_alm(fn_b, blm, overwrite=True)
        return elm if which == 'e' else blm

    def get_sim_elm(self, idx):
        """Inverse-variance filtered E-mode alm."""
        return self._get_sim_eblm(idx, 'e')

    def get_sim_blm(self, idx):
        """Inverse-variance filtered B-mode alm."""
        return self._get_sim_eblm(idx, 'b')

    def get_sim_tmliklm(self, idx):
        """Wiener-filtered temperature alm."""
        return hp.almxfl(self.get_sim_tlm(idx), self.cl['tt'])

    def get_sim_emliklm(self, idx):
        return hp.almxfl(self.get_sim_elm(idx), self.cl['ee'])

    def get_sim_bmliklm(self, idx):
        return hp.almxfl(self.get_sim_blm(idx), self.cl['bb'])

    # -- the same products as CUDA tensors (no reference counterpart: there everything is a numpy array on the host).
    #    `qest.library` asks for these when both of its filtering libraries provide them, which keeps a simulation on the
    #    GPU from the simulated map to the quadratic estimate: maps from `sim_lib.get_sim_*map_dev`, the filter through
    #    `_apply_ivf_*_dev`, and the last few filtered skies in a small device cache.  Disk caching (`self.cache`, same file
    #    names as above) happens on a worker thread from pinned host copies, off the GPU's critical path.
    _DEV_CACHE_SIMS = 4

    def _dev_store(self):
        if not hasattr(self, '_dev_alms'):
            import collections
            self._dev_alms = collections.OrderedDict()
        return self._dev_alms

    def _sim_map_dev(self, idx, which):
        fun = getattr(self.sim_lib, 'get_sim_%smap_dev' % which, None)
        if fun is not None:
            return fun(idx)
        return getattr(self.sim_lib, 'get_sim_%smap' % which)(idx)          # numpy (or tensors with device_maps)

    def _write_async(self, fn, tensor):
        """tensor -> pinned host copy on a side stream -> hp.write_alm on a worker thread (write-then-rename).
        Pinned buffers are pooled (cudaHostAlloc of tens of MB costs milliseconds) and the copy waits on an event of the
        compute stream, so neither the allocation nor the D2H transfer sits on the GPU's critical path."""
        import concurrent.futures as cf
        import torch
        if not hasattr(self, '_io_pool'):
            self._io_pool = cf.ThreadPoolExecutor(max_workers=2)
            self._io_pending = []
            self._io_stream = torch.cuda.Stream()
            self._io_free = {}
        n = tensor.numel()
        free = self._io_free.setdefault(n, [])
        host = free.pop() if free else torch.empty(n, dtype=torch.complex128, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record()                                        # the tensor is complete on the compute stream
        with torch.cuda.stream(self._io_stream):
            self._io_stream.wait_event(ready)
            host.copy_(tensor, non_blocking=True)
            tensor.record_stream(self._io_stream)             # keep the device memory alive until the copy has run
            done = torch.cuda.Event()
            done.record()

        def job():
            done.synchronize()
            root, ext = os.path.splitext(fn)
            tmp = '%s.tmp%d%s' % (root, os.getpid(), ext)        # keeps the extension hp.write_alm dispatches on
            hp.write_alm(tmp, host.numpy(), overwrite=True)
            os.replace(tmp, fn)
            free.append(host)
        self._io_pending = [f for f in self._io_pending if not f.done()]
        self._io_pending.append(self._io_pool.submit(job))

    def flush(self):
        """waits for the asynchronous cache writes of the device accessors"""
        for f in getattr(self, '_io_pending', []):
            f.result()
        self._io_pending = []

    # T and P filters that are worth running side by side (two host threads, two streams): set by library_cinv_sepTP,
    # whose conjugate-gradient solves each leave much of the GPU idle in their multigrid preconditioners
    _TP_CONCURRENT = False

    def _filter_t_dev(self, idx, tmap=None):
        from .. import sht
        soltn = None if self.soltn_lib is None else self.soltn_lib.get_sim_tmliklm(idx)
        if hasattr(self, '_apply_ivf_t_dev'):
            return self._apply_ivf_t_dev(self._sim_map_dev(idx, 't') if tmap is None else tmap, soltn=soltn)
        return sht.dev_alm(self._apply_ivf_t(self.sim_lib.get_sim_tmap(idx), soltn=soltn))

    def _filter_p_dev(self, idx, pmap=None):
        from .. import sht
        soltn = None
        if self.soltn_lib is not None:
            soltn = np.array([self.soltn_lib.get_sim_emliklm(idx), self.soltn_lib.get_sim_bmliklm(idx)])
        if hasattr(self, '_apply_ivf_p_dev'):
            return self._apply_ivf_p_dev(self._sim_map_dev(idx, 'p') if pmap is None else pmap, soltn=soltn)
        e, b = self._apply_ivf_p(self.sim_lib.get_sim_pmap(idx), soltn=soltn)
        return sht.dev_alm(e), sht.dev_alm(b)

    # ---- the T and the P filter as two lanes (one worker thread and one stream each; `sht.use_lane` keeps their plans and
    #      reduction scratch apart).  `_submit_tp` makes the simulated maps on the calling stream and queues the two solves;
    #      `_collect_tp` joins them.  With `prefetch_dev` the lanes work ahead of the caller: while it evaluates the
    #      estimator of simulation i on its own stream, the lanes filter simulations i + 1, i + 2 -- the polarization lane,
    #      whose solve is the shorter one, runs ahead and fills the SMs the temperature lane leaves idle.
    def _lanes(self):
        import concurrent.futures as cf
        import torch
        if not hasattr(self, '_t_pool'):
            self._t_pool = cf.ThreadPoolExecutor(max_workers=1)
            self._p_pool = cf.ThreadPoolExecutor(max_workers=1)
            # The temperature solve is the longer chain (13-16 top-level iterations against 5 at Planck-like noise): its
            # stream outranks the polarization solve's and the caller's
            self._t_stream = torch.cuda.Stream(priority=int(os.environ.get('PLK_TP_TPRIO', '-2')))
            self._p_stream = torch.cuda.Stream(priority=0)
            self._pending = {}

    def _tp_enabled(self):
        return self._TP_CONCURRENT and os.environ.get('PLK_TP_CONCURRENT', '1') != '0' \
            and getattr(self, '_tp_ready', lambda: True)()

    def _submit_tp(self, idx):
        import torch
        self._lanes()
        tmap, pmap = self._sim_map_dev(idx, 't'), self._sim_map_dev(idx, 'p')
        dev = torch.cuda.current_device()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        for m, st in [(tmap, self._t_stream)] + [(x, self._p_stream) for x in pmap]:
            if isinstance(m, torch.Tensor):
                m.record_stream(st)                            # made on the calling stream, read on the lane's

        def job(stream, fn, arg):
            def work():
                torch.cuda.set_device(dev)
                with torch.cuda.stream(stream):
                    stream.wait_event(ready)
                    out = fn(idx, arg)
                    done = torch.cuda.Event()
                    done.record(stream)
                return out, done
            return work
        self._pending[idx] = (self._t_pool.submit(job(self._t_stream, self._filter_t_dev, tmap)),
                              self._p_pool.submit(job(self._p_stream, self._filter_p_dev, pmap)))

    def _collect_tp(self, idx):
        import torch
        ft, fp = self._pending.pop(idx)
        try:
            t, done_t = ft.result()                            # re-raises what the lane raised
        finally:
            (e, b), done_p = fp.result()
        main = torch.cuda.current_stream()
        main.wait_event(done_t)
        main.wait_event(done_p)
        for x in (t, e, b):
            x.record_stream(main)                              # allocated on the lanes' streams, used on this one
        return t, e, b

    def _needs(self, idx):
        """(T to filter, P to filter): not in the device store, not on disk"""
        ent = self._dev_store().get(idx, {})
        fn_t, fn_e, fn_b = (self._fname(idx, f) for f in 'teb')
        return ('t' not in ent and not os.path.exists(fn_t)), \
               ('e' not in ent and not (os.path.exists(fn_e) and os.path.exists(fn_b)))

    def prefetch_dev(self, idxs):
        """Hint: the simulations `idxs` will be asked for next (`get_sim_teblm_dev`), in this order.  Their maps are
        simulated now, on the calling stream, and their filters queued on the two lanes; returns at once.  No effect
        for libraries whose filters do not run side by side, before both chains are warm, or for what is cached."""
        if not self._tp_enabled():
            return
        for idx in idxs:
            if idx not in getattr(self, '_pending', {}) and all(self._needs(idx)):
                self._submit_tp(idx)

    def get_sim_teblm_dev(self, idx, fields='teb'):
        """Inverse-variance filtered alms of simulation idx as complex128 CUDA tensors, in the order of `fields`."""
        from .. import sht
        store = self._dev_store()
        ent = store.setdefault(idx, {})
        store.move_to_end(idx)
        fn_t, fn_e, fn_b = (self._fname(idx, f) for f in 'teb')
        need_t = 't' in fields and 't' not in ent
        need_p = ('e' in fields or 'b' in fields) and 'e' not in ent
        if need_t and os.path.exists(fn_t):
            ent['t'], need_t = sht.dev_alm(hp.read_alm(fn_t)), False
        if need_p and os.path.exists(fn_e) and os.path.exists(fn_b):
            ent['e'], ent['b'], need_p = sht.dev_alm(hp.read_alm(fn_e)), sht.dev_alm(hp.read_alm(fn_b)), False
        if idx in getattr(self, '_pending', {}):
            ent['t'], ent['e'], ent['b'] = self._collect_tp(idx)      # prefetched: both filters ran, whatever `fields` asks
            need_t = need_p = True
        elif need_t and need_p and self._tp_enabled():
            self._submit_tp(idx)
            ent['t'], ent['e'], ent['b'] = self._collect_tp(idx)
        else:
            if need_t:
                ent['t'] = self._filter_t_dev(idx)
            if need_p:
                ent['e'], ent['b'] = self._filter_p_dev(idx)
        if self.cache:
            for f, need in (('t', need_t), ('e', need_p), ('b', need_p)):
                if need:
                    self._write_async(self._fname(idx, f), ent[f])
        while len(store) > self._DEV_CACHE_SIMS:
            store.popitem(last=False)
        return tuple(ent[f] for f in fields)

    def _cl_dev(self, key, lmax):
        from .. import sht
        if not hasattr(self, '_cl_d'):
            self._cl_d = {}
        if (key, lmax) not in self._cl_d:
            self._cl_d[(key, lmax)] = sht.dev_fl(self.cl[key], lmax)
        return self._cl_d[(key, lmax)]

    def get_sim_mliklm_dev(self, idx, fields='teb'):
        """Wiener-filtered alms C_l^{aa} x (inverse-variance filtered a) on the device"""
        from .. import sht
        out = []
        for f, a in zip(fields, self.get_sim_teblm_dev(idx, fields)):
            out.append(sht.almxfl(a, self._cl_dev(f + f, sht.alm_lmax(a.numel()))))
        return tuple(out)


class library_jTP(object):
    """Inverse-variance and Wiener filtering of a simulation library, T and P filtered JOINTLY
    (reference: filt_simple.py:187-343).  Concrete filters provide `hashdict`, `get_fmask`, `_apply_ivf`, `get_fal`."""

    def __init__(self, lib_dir, sim_lib, cl_weights, soltn_lib=None, cache=True):
        assert np.all([k in cl_weights.keys() for k in ['tt', 'ee', 'bb']])
        self.lib_dir = lib_dir
        self.sim_lib = sim_lib
        self.cl = cl_weights
        self.soltn_lib = soltn_lib
        self.cache = cache
        fn_hash = os.path.join(lib_dir, 'filt_hash.pk')
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            if not os.path.exists(fn_hash):
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fn_hash, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=fn_hash)

    def hashdict(self):
        assert 0, 'override this'

    def get_fmask(self):
        assert 0, 'override this'

    def _apply_ivf(self, tqumap, soltn=None):
        assert 0, 'override this'

    def get_fal(self):
        r"""Isotropic matrix approximation :math:`F_\ell \sim (C_\ell + N_\ell / b_\ell^2)^{-1}` ('tt', 'ee', 'te', ...)."""
        assert 0, 'override this'

    def _get_alms(self, a, idx):
        assert a in ['t', 'e', 'b']
        tfname = os.path.join(self.lib_dir, 'sim_%04d_tlm.fits' % idx if idx >= 0 else 'dat_tlm.fits')
        fname = tfname.replace('tlm.fits', a + 'lm.fits')
        if not os.path.exists(fname):
            T = self.sim_lib.get_sim_tmap(idx)
            Q, U = self.sim_lib.get_sim_pmap(idx)
            soltn = None
            if self.soltn_lib is not None:
                soltn = (self.soltn_lib.get_sim_tmliklm(idx), self.soltn_lib.get_sim_emliklm(idx),
                         self.soltn_lib.get_sim_bmliklm(idx))
            tlm, elm, blm = self._apply_ivf([T, Q, U], soltn=soltn)
            if not self.cache:
                return {'t': tlm, 'e': elm, 'b': blm}[a]
            hp.write_alm(tfname, tlm, overwrite=True)
            hp.write_alm(tfname.replace('tlm.fits', 'elm.fits'), elm, overwrite=True)
            hp.write_alm(tfname.replace('tlm.fits', 'blm.fits'), blm, overwrite=True)
        return hp.read_alm(fname)

    def get_sim_tlm(self, idx):
        return self._get_alms('t', idx)

    def get_sim_elm(self, idx):
        return self._get_alms('e', idx)

    def get_sim_blm(self, idx):
        return self._get_alms('b', idx)

    def _mlik(self, a, idx):
        """Wiener-filtered alm: sum_b C_l^{ab} (inverse-variance filtered b), b over t, e, b (reference: :294-343)."""
        ret = hp.almxfl(self._get_alms(a, idx), self.cl[a + a])
        for b in 'teb':
            if b == a:
                continue
            cl = self.cl.get(a + b, self.cl.get(b + a, None))
            if cl is not None:
                ret = ret + hp.almxfl(self._get_alms(b, idx), cl)
        return ret

    def get_sim_tmliklm(self, idx):
        return self._mlik('t', idx)

    def get_sim_emliklm(self, idx):
        return self._mlik('e', idx)

    def get_sim_bmliklm(self, idx):
        return self._mlik('b', idx)


class library_fullsky_sepTP(library_sepTP):
    """Full-sky isotropic filter: filtered alm = f_l / transf_l * map2alm(map) (reference: filt_simple.py:346-407).

    Args:
        lib_dir, sim_lib: as above
        nside: resolution of the simulation library
        transf: transfer function (array, or dict with 't', 'e', 'b')
        cl_len: spectra for the Wiener-filtered alms
        ftl, fel, fbl: isotropic filters
    """

    def __init__(self, lib_dir, sim_lib, nside, transf, cl_len, ftl, fel, fbl, cache=False):
        transfd = transf if isinstance(transf, dict) else {'t': transf, 'e': transf, 'b': transf}
        assert all(k in transfd for k in 'teb')
        self.sim_lib = sim_lib
        self.ftl, self.fel, self.fbl = ftl, fel, fbl
        self.lmax_fl = max(len(ftl), len(fel), len(fbl)) - 1
        self.nside = nside
        self.transf = transfd
        super(library_fullsky_sepTP, self).__init__(lib_dir, sim_lib, cl_len, cache=cache)

    def hashdict(self):
        return {'sim_lib': self.sim_lib.hashdict(), 'transf': utils.clhash(self.transf['t']),
                'cl_len': {k: utils.clhash(self.cl[k]) for k in ['tt', 'ee', 'bb']},
                'ftl': utils.clhash(self.ftl), 'fel': utils.clhash(self.fel), 'fbl': utils.clhash(self.fbl)}

    def get_fmask(self):
        return np.ones(hp.nside2npix(self.nside), dtype=float)

    def get_tal(self, a):
        assert a.lower() in ['t', 'e', 'b']
        return utils.cli(self.transf[a.lower()])

    def get_ftl(self):
        return np.copy(self.ftl)

    def get_fel(self):
        return np.copy(self.fel)

    def get_fbl(self):
        return np.copy(self.fbl)

    def _apply_ivf_t(self, tmap, soltn=None):
        assert len(tmap) == hp.nside2npix(self.nside), (hp.npix2nside(tmap.size), self.nside)
        alm = hp.map2alm(tmap, lmax=self.lmax_fl, iter=0)
        return hp.almxfl(alm, self.get_ftl() * utils.cli(self.transf['t'][:len(self.ftl)]))

    def _apply_ivf_p(self, pmap, soltn=None):
        assert len(pmap[0]) == hp.nside2npix(self.nside) and len(pmap[0]) == len(pmap[1])
        elm, blm = hp.map2alm_spin([m for m in pmap], 2, lmax=self.lmax_fl)
        elm = hp.almxfl(elm, self.get_fel() * utils.cli(self.transf['e'][:len(self.fel)]))
        blm = hp.almxfl(blm, self.get_fbl() * utils.cli(self.transf['b'][:len(self.fbl)]))
        return elm, blm


    # device-resident forms: one analysis with the per-l filter fused (maps may be numpy arrays or CUDA tensors)
    def _ivf_fl_dev(self, a):
        from .. import sht
        if not hasattr(self, '_fl_d'):
            self._fl_d = {}
        if a not in self._fl_d:
            fl = {'t': self.get_ftl, 'e': self.get_fel, 'b': self.get_fbl}[a]()
            self._fl_d[a] = sht.dev_fl(fl * utils.cli(self.transf[a][:len(fl)]), self.lmax_fl)
        return self._fl_d[a]

    def _apply_ivf_t_dev(self, tmap, soltn=None):
        from .. import sht
        assert len(tmap) == hp.nside2npix(self.nside), (len(tmap), self.nside)
        return sht.get_plan(self.nside, self.lmax_fl).map2alm(sht.dev_map(tmap), fl=self._ivf_fl_dev('t'))

    def _apply_ivf_p_dev(self, pmap, soltn=None):
        from .. import sht
        assert len(pmap[0]) == hp.nside2npix(self.nside) and len(pmap[0]) == len(pmap[1])
        return sht.get_plan(self.nside, self.lmax_fl).map2alm_spin(sht.dev_map(pmap[0]), sht.dev_map(pmap[1]), 2,
                                                                   flg=self._ivf_fl_dev('e'), flc=self._ivf_fl_dev('b'))


class library_fullsky_alms_sepTP(library_fullsky_sepTP):
    """The same isotropic filter for a simulation library that hands out harmonic coefficients: its `get_sim_tmap`
    returns tlm and `get_sim_pmap` returns (elm, blm), so no transform is involved (reference: filt_simple.py:409-470)."""

    def __init__(self, lib_dir, sim_lib, transf, cl_len, ftl, fel, fbl, cache=False):
        super(library_fullsky_alms_sepTP, self).__init__(lib_dir, sim_lib, None, transf, cl_len, ftl, fel, fbl, cache=cache)

    def get_fmask(self):
        return np.array([1.])       # for compatibility only

    def _apply_ivf_t(self, tlm, soltn=None):
        return hp.almxfl(tlm, self.get_ftl() * utils.cli(self.transf['t'][:len(self.ftl)]))

    def _apply_ivf_p(self, eblm, soltn=None):
        elm = hp.almxfl(eblm[0], self.get_fel() * utils.cli(self.transf['e'][:len(self.fel)]))
        blm = hp.almxfl(eblm[1], self.get_fbl() * utils.cli(self.transf['b'][:len(self.fbl)]))
        return elm, blm


class library_apo_sepTP(library_sepTP):
    """Apodised-mask + isotropic filter (reference: filt_simple.py:473-534)."""

    def __init__(self, lib_dir, sim_lib, apomask_path, cl_len, transf, ftl, fel, fbl, cache=False):
        assert len(transf) >= max(len(ftl), len(fel), len(fbl))
        assert os.path.exists(apomask_path)
        self.ftl, self.fel, self.fbl = ftl, fel, fbl
        self.transf = transf
        self.lmax_fl = max(len(ftl), len(fel), len(fbl)) - 1
        self.apomask_path = apomask_path
        self.nside = hp.npix2nside(hp.read_map(apomask_path).size)
        super(library_apo_sepTP, self).__init__(lib_dir, sim_lib, cl_len, cache=cache)

    def hashdict(self):
        return {'sim_lib': self.sim_lib.hashdict(), 'apomask': self.apomask_path, 'transf': utils.clhash(self.transf),
                'cl_len': {k: utils.clhash(self.cl[k]) for k in ['tt', 'ee', 'bb']},
                'ftl': utils.clhash(self.ftl), 'fel': utils.clhash(self.fel), 'fbl': utils.clhash(self.fbl)}

    def get_fmask(self):
        return hp.read_map(self.apomask_path)

    def get_tal(self, a):
        assert a.lower() in ['t', 'e', 'b']
        return utils.cli(self.transf)

    def get_ftl(self):
        return np.copy(self.ftl)

    def get_fel(self):
        return np.copy(self.fel)

    def get_fbl(self):
        return np.copy(self.fbl)

    def _apply_ivf_t(self, tmap, soltn=None):
        alm = hp.map2alm(tmap * self.get_fmask(), lmax=self.lmax_fl, iter=0)
        return hp.almxfl(alm, self.get_ftl() * utils.cli(self.transf[:len(self.ftl)]))

    def _apply_ivf_p(self, pmap, soltn=None):
        mask = self.get_fmask()
        elm, blm = hp.map2alm_spin([m * mask for m in pmap], 2, lmax=self.lmax_fl)
        elm = hp.almxfl(elm, self.get_fel() * utils.cli(self.transf[:len(self.fel)]))
        blm = hp.almxfl(blm, self.get_fbl() * utils.cli(self.transf[:len(self.fbl)]))
        return elm, blm
