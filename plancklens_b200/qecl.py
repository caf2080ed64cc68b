"""QE (cross-)power spectra library (reference: plancklens/qecl.py:12-148).

Combines two `qest.library` instances and a set of mean-field simulations into raw spectra
:math:`\\frac{1}{(2L+1) f_{sky}} \\sum_M \\hat\\phi^A_{LM} \\hat\\phi^{B\\dagger}_{LM}` after mean-field subtraction.  The
subtraction and `alm2cl` run on the GPU (`plk_alm_axpy_dev`, `plk_alm2cl_dev`); spectra are cached in the reference's
sqlite `npdb` (`cldb.db`) under the reference's key strings, so caches are interchangeable.
"""
import os
import pickle as pk

import numpy as np

from . import sht, utils
from .helpers import mpi, sql


class library(object):
    def __init__(self, lib_dir, qeA, qeB, mc_sims_mf):
        self.lib_dir = lib_dir
        self.prefix = lib_dir
        self.qeA = qeA
        self.qeB = qeB
        self.mc_sims_mf = np.asarray(mc_sims_mf, dtype=int)
        fsname = os.path.join(lib_dir, 'fskies.dat')
        hname = os.path.join(self.lib_dir, 'qcl_sim_hash.pk')
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            if not os.path.exists(fsname):
                ms = {1: self.qeA.get_mask(1), 2: self.qeA.get_mask(2), 3: self.qeB.get_mask(1), 4: self.qeB.get_mask(2)}
                assert np.all([m.shape == ms[1].shape for m in ms.values()])
                fskies = {}
                for i in [1, 2, 3, 4]:
                    for j in [1, 2, 3, 4][i - 1:]:
                        fskies[10 * i + j] = np.mean(ms[i] * ms[j])
                fskies[1234] = np.mean(ms[1] * ms[2] * ms[3] * ms[4])
                with open(fsname, 'w') as f:
                    for lab in np.sort(list(fskies.keys())):
                        f.write('%4s %.5f \n' % (lab, fskies[lab]))
            if not os.path.exists(hname):
                with open(hname, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(hname, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=hname)
        self.npdb = sql.npdb(os.path.join(lib_dir, 'cldb.db'))     # the reference's cache, same keys (qecl.py:61)
        fskies = {}
        with open(fsname) as f:
            for line in f:
                key, val = line.split()
                fskies[int(key)] = float(val)
        self.fskies = fskies
        self.fsky1234, self.fsky11, self.fsky12, self.fsky22 = fskies[1234], fskies[11], fskies[12], fskies[22]

    def hashdict(self):
        return {'qeA': self.qeA.hashdict(), 'qeB': self.qeB.hashdict(), 'mc_sims_mf': self._mcmf_hash()}

    def _mcmf_hash(self):
        return utils.mchash(self.mc_sims_mf)

    def get_lmaxqcl(self, k1, k2):
        return min(self.qeA.get_lmax_qlm(k1), self.qeB.get_lmax_qlm(k2))

    def load_sim_qcl(self, k1, idx, k2=None, lmax=None):
        """Same as get_sim_qcl without triggering its calculation"""
        return self.get_sim_qcl(k1, idx, k2=k2, lmax=lmax, calc=False)

    def get_sim_qcl(self, k1, idx, k2=None, lmax=None, recache=False, calc=True):
        """QE (cross-)power spectrum of simulation idx (idx = -1: data), mean field subtracted, / fsky."""
        if k2 is None:
            k2 = k1
        assert k1 in self.qeA.keys and k2 in self.qeB.keys, (k1, k2)
        assert idx not in self.mc_sims_mf, idx
        lmax_qcl = self.get_lmaxqcl(k1, k2)
        lmax_out = lmax or lmax_qcl
        assert lmax_out <= lmax_qcl
        tag = '%04d' % idx if idx >= 0 else 'dat'
        assert idx >= -1
        # key strings and recache semantics of the reference (qecl.py:101-119): spectra live in the sqlite npdb
        # `cldb.db`, so a lib_dir written by either code serves the other
        fname = os.path.join(self.lib_dir, 'sim_qcl_k1%s_k2%s_lmax%s_%s_%s.dat' % (k1, k2, lmax_qcl, tag, self._mcmf_hash()))
        if calc:
            recache = False
        if calc and (self.npdb.get(fname) is None or recache):
            qlmA = sht.dev_alm(self.qeA.get_sim_qlm(k1, idx, lmax=lmax_qcl))
            if (k1 == k2) and (self.qeA is self.qeB):
                qlmB = qlmA.clone()
            else:
                qlmB = sht.dev_alm(self.qeB.get_sim_qlm(k2, idx, lmax=lmax_qcl))
            sht.alm_axpy(qlmA, sht.dev_alm(self.qeA.get_sim_qlm_mf(k1, self.mc_sims_mf[0::2], lmax=lmax_qcl)), -1.0)
            sht.alm_axpy(qlmB, sht.dev_alm(self.qeB.get_sim_qlm_mf(k2, self.mc_sims_mf[1::2], lmax=lmax_qcl)), -1.0)
            if recache and self.npdb.get(fname) is not None:
                self.npdb.remove(fname)
            self.npdb.add(fname, self._alm2clfsky1234(qlmA, qlmB, k1, k2))
        return self.npdb.get(fname)[:lmax_out + 1] / self.fskies[1234]

    def get_dat_qcl(self, k1, k2=None, lmax=None, recache=False):
        """QE (cross-)power spectrum of the data maps (index -1); `qecl.average.get_dat_qcl` calls it, although the
        reference's `library` does not define it (qecl.py:199 raises AttributeError there)"""
        return self.get_sim_qcl(k1, -1, k2=k2, lmax=lmax, recache=recache)

    def get_sim_stats_qcl(self, k1, mc_sims, k2=None, recache=False):
        """Mean and scatter of the QE spectra over mc_sims, as a cached `utils.stats` instance (reference: qecl.py:126-145)."""
        if k2 is None:
            k2 = k1
        tfname = os.path.join(self.lib_dir, 'sim_qcl_stats_%s_%s_%s.pk' % (k1, k2, utils.mchash(mc_sims)))
        if not os.path.exists(tfname) or recache:
            st = utils.stats(self.get_lmaxqcl(k1, k2) + 1, docov=False)
            for idx in mc_sims:
                st.add(self.get_sim_qcl(k1, idx, k2=k2))
            with open(tfname, 'wb') as f:
                pk.dump(st, f, protocol=2)
        with open(tfname, 'rb') as f:
            return pk.load(f)

    def _alm2clfsky1234(self, qlm1, qlm2, k1, k2):
        return sht.alm2cl(qlm1, qlm2).cpu().numpy()


class average:
    """Average of several QE spectra libraries (reference: qecl.py:151-223).

        Args:
            lib_dir: the statistics are cached there
            qcls_lib: list of `qecl.library` instances
    """

    def __init__(self, lib_dir, qcls_lib):
        self.lib_dir = lib_dir
        self.qclibs = qcls_lib
        hname = os.path.join(lib_dir, 'qeclav_hash.pk')
        if mpi.rank == 0:
            os.makedirs(lib_dir, exist_ok=True)
            if not os.path.exists(hname):
                with open(hname, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(hname, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=hname)
        self.mc_sims_mf = np.sort(np.unique(np.concatenate([q.mc_sims_mf for q in self.qclibs])))

    def hashdict(self):
        return {'qcl_lib %s' % i: q.hashdict() for i, q in enumerate(self.qclibs)}

    def get_lmaxqcl(self, k1, k2):
        return np.min([q.get_lmaxqcl(k1, k2) for q in self.qclibs])

    def _mean(self, get):
        return sum(get(q) for q in self.qclibs) / len(self.qclibs)

    def get_sim_qcl(self, k1, idx, k2=None, lmax=None):
        if lmax is None:
            lmax = self.get_lmaxqcl(k1, k2)
        return self._mean(lambda q: q.get_sim_qcl(k1, idx, k2=k2, lmax=lmax))

    def get_dat_qcl(self, k1, k2=None, lmax=None):
        if lmax is None:
            lmax = self.get_lmaxqcl(k1, k2)
        return self._mean(lambda q: q.get_dat_qcl(k1, k2=k2, lmax=lmax))

    def get_sim_stats_qcl(self, k1, mc_sims, k2=None, recache=False, lmax=None):
        if k2 is None:
            k2 = k1
        if lmax is None:
            lmax = self.get_lmaxqcl(k1, k2)
        tfname = os.path.join(self.lib_dir, 'sim_qcl_stats_%s_%s_%s_%s.pk' % (k1, k2, lmax, utils.mchash(mc_sims)))
        if not os.path.exists(tfname) or recache:
            st = utils.stats(lmax + 1, docov=False)
            for idx in mc_sims:
                st.add(self.get_sim_qcl(k1, idx, k2=k2, lmax=lmax))
            with open(tfname, 'wb') as f:
                pk.dump(st, f, protocol=2)
        with open(tfname, 'rb') as f:
            return pk.load(f)
