r""":math:`N^{(1)}_L` bias library: the surface parameter files construct (reference: plancklens/n1/n1.py).

The reference evaluates :math:`N^{(1)}` with flat-sky integrals in a Fortran extension (`n1f.f90`); that integral is
outside the spherical-harmonic hot path this package implements and there is **no backend for it
here**: `HASN1F` is False, exactly as in a reference install whose extension failed to build (n1.py:28-32).  What is
mirrored is everything around it, so that parameter files import and caches are shared:

* `library_n1(lib_dir, cltt, clte, clee, lmaxphi, dL, lps)`: same default `lps` nodes, hash file, and the two sqlite
  caches `npdb.db` (splined curves) and `fldb.db` (single multipoles) under the reference's key strings;
* `get_n1`: key ordering, decomposition of the derived estimators ('p', 'p_p', 'p_tp', 'p_eb', ...) into fundamental
  pairs with their weights, splining of the sampled multipoles, cache look-ups, `recache` / `remove_only`.  Curves present
  in the cache (e.g. written by a reference install sharing the directory) are served; a missing multipole raises
  `NotImplementedError` where the reference would call `n1f.n1l`, unless an integrator with that argument list has been
  attached as `library_n1.n1l` (how the tests pin this layer to the reference, which is given the same stand-in).
"""
import os
import pickle as pk

import numpy as np

from ..helpers import mpi, sql
from ..utils import clhash, cli, hash_check

HASN1F = False

estimator_keys = ['ptt', 'pte', 'pet', 'pee', 'peb', 'pbe', 'ptb', 'pbt',
                  'xtt', 'xte', 'xet', 'xee', 'xeb', 'xbe', 'xtb', 'xbt',
                  'stt', 'ftt']
estimator_keys_derived = ['p', 'p_p', 'p_tp', 'p_eb', 'p_te', 'p_tb',
                          'f', 'f_p', 'f_tp', 'f_eb', 'f_te', 'f_tb',
                          'x', 'x_p', 'x_tp', 'x_eb', 'x_te', 'x_tb']


def _get_est_derived(k, lmax):
    """[(fundamental key, weight array)] of a derived estimator (reference: n1.py:50-82)."""
    one = np.ones(lmax + 1, dtype=float)
    g = k[0]
    if k in ['p', 'x', 'f']:
        return [(g + 'tt', one), (g + 'te', 2. * one), (g + 'tb', 2. * one), (g + 'ee', one), (g + 'eb', 2. * one)]
    if k in ['p_tp', 'x_tp', 'f_tp']:
        return [(g + 'tt', one), (g + 'ee', one), (g + 'eb', 2. * one)]
    if k in ['p_p', 'x_p', 'f_p']:
        return [(g + 'ee', one), (g + 'eb', 2. * one)]
    if k in ['p_te', 'x_te', 'p_tb', 'x_tb', 'p_eb', 'x_eb']:
        return [(k.replace('_', ''), 2. * one)]
    if k in estimator_keys:
        return [k, one]        # as in the reference (a flat pair, not a list of pairs)
    assert 0, k


def _default_lps(lmaxphi):
    """integration nodes in the anisotropy-source multipole (reference: n1.py:104-117)"""
    lps = [1] + list(range(2, 111, 10))
    lps += list(range(lps[-1] + 30, 580, 30))
    lps += list(range(lps[-1] + 100, lmaxphi // 2, 100))
    lps += list(range(lps[-1] + 300, lmaxphi, 300))
    if lps[-1] != lmaxphi:
        lps.append(lmaxphi)
    return np.array(lps)


def _sample_Ls(Lmax):
    return np.unique(np.concatenate([np.arange(1, 11), np.arange(1, Lmax + 1)[::20], [Lmax]]))


class library_n1:
    r"""N1 bias library (reference: n1.py:89-316).

        Args:
            lib_dir: results are cached there
            cltt, clte, clee: CMB spectra of the maps (and, by default, of the QE weights)
            lmaxphi: maximum multipole of the anisotropy source
            dL, lps: flat-sky integration parameters (kept for the hash and the cache keys)
    """

    def __init__(self, lib_dir, cltt, clte, clee, lmaxphi=2500, dL=10, lps=None):
        self.dL = dL
        self.lps = _default_lps(lmaxphi) if lps is None else lps
        self.cltt, self.clte, self.clee = cltt, clte, clee
        self.lmaxphi = self.lps[-1]
        self.n1 = {}
        self.n1l = None      # hook: a callable with the argument list of the reference's `n1f.n1l` (n1.py:305-308)
        fn_hash = os.path.join(lib_dir, 'n1_hash.pk')
        if mpi.rank == 0:
            os.makedirs(lib_dir, exist_ok=True)
            if not os.path.exists(fn_hash):
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fn_hash, 'rb') as f:
            hash_check(self.hashdict(), pk.load(f), fn=fn_hash)
        self.npdb = sql.npdb(os.path.join(lib_dir, 'npdb.db'))
        self.fldb = sql.fldb(os.path.join(lib_dir, 'fldb.db'))
        self.lib_dir = lib_dir

    def hashdict(self):
        return {'cltt': clhash(self.cltt), 'clte': clhash(self.clte), 'clee': clhash(self.clee), 'dL': self.dL, 'lps': self.lps}

    @staticmethod
    def _key(head, kA, kB, k_ind, cl_kind, fals, clfids, tail=''):
        idx = head + 'kA' + kA + '_kB' + kB + '_ind' + k_ind + '_clpp' + clhash(cl_kind)
        for name, fl in zip(('ftlA', 'felA', 'fblA', 'ftlB', 'felB', 'fblB'), fals):
            idx += '_' + name + clhash(fl)
        for name, cl in zip(('clttfid', 'cltefid', 'cleefid'), clfids):
            idx += '_' + name + clhash(cl)
        return idx + tail

    def get_n1(self, kA, k_ind, cl_kind, ftlA, felA, fblA, Lmax, kB=None, ftlB=None, felB=None, fblB=None,
               clttfid=None, cltefid=None, cleefid=None, n1_flat=lambda ell: np.ones(len(ell), dtype=float),
               recache=False, remove_only=False, sglLmode=True):
        r"""N1 bias of the spectrum of estimators kA, kB for anisotropy source `k_ind` with spectrum `cl_kind`, up to Lmax
        (reference: n1.py:141-271; same arguments)."""
        if kB is None:
            kB = kA
        if kA[0] == 's' or kB[0] == 's':
            assert kA[0] == kB[0], 'point source implemented following the gradient convention: pick a sign first'
        ftlB = ftlA if ftlB is None else ftlB
        felB = felA if felB is None else felB
        fblB = fblA if fblB is None else fblB
        clttfid = self.cltt if clttfid is None else clttfid
        cltefid = self.clte if cltefid is None else cltefid
        cleefid = self.clee if cleefid is None else cleefid
        fid = dict(clttfid=clttfid, cltefid=cltefid, cleefid=cleefid, n1_flat=n1_flat, sglLmode=sglLmode)

        if kA in estimator_keys and kB in estimator_keys:
            if kA < kB:       # canonical order of the pair: legs swapped
                return self.get_n1(kB, k_ind, cl_kind, ftlB, felB, fblB, Lmax, ftlB=ftlA, felB=felA, fblB=fblA, kB=kA, **fid)
            fals, clfids = (ftlA, felA, fblA, ftlB, felB, fblB), (clttfid, cltefid, cleefid)
            idx = self._key('splined_', kA, kB, k_ind, cl_kind, fals, clfids, '_Lmax%s' % Lmax)
            ret = self.npdb.get(idx)
            if ret is not None:
                if not recache and not remove_only:
                    return ret
                self.npdb.remove(idx)
                if remove_only:
                    return np.zeros_like(ret)
            Ls = _sample_Ls(Lmax)
            assert sglLmode, 'the vectorised Fortran call has no counterpart here'
            n1L = np.zeros(len(Ls), dtype=float)
            for i, L in enumerate(Ls[mpi.rank::mpi.size]):
                n1L[i] = self._get_n1_L(L, kA, kB, k_ind, cl_kind, *fals, *clfids, remove_only=remove_only)
            if mpi.size > 1:      # every rank reloads what the others cached
                mpi.barrier()
                for i, L in enumerate(Ls):
                    n1L[i] = self._get_n1_L(L, kA, kB, k_ind, cl_kind, *fals, *clfids, remove_only=remove_only)
                mpi.barrier()
            from scipy.interpolate import UnivariateSpline as spline
            ret = np.zeros(Lmax + 1)
            ell = np.arange(1, Lmax + 1) * 1.
            ret[1:] = spline(Ls, n1L * n1_flat(Ls), s=0., ext='raise', k=3)(ell) * cli(n1_flat(ell))
            self.npdb.add(idx, ret)
            return ret

        termsA = _get_est_derived(kA, Lmax) if kA in estimator_keys_derived else None
        termsB = _get_est_derived(kB, Lmax) if kB in estimator_keys_derived else None
        assert (termsA is not None or kA in estimator_keys) and (termsB is not None or kB in estimator_keys), (kA, kB)
        one = np.ones(Lmax + 1)
        ret = 0.
        for tk1, cl1 in (termsA if termsA is not None else [(kA, one)]):
            for tk2, cl2 in (termsB if termsB is not None else [(kB, one)]):
                tret = self.get_n1(tk1, k_ind, cl_kind, ftlA, felA, fblA, Lmax, ftlB=ftlB, felB=felB, fblB=fblB, kB=tk2, **fid)
                ret = ret + tret * cl1[:Lmax + 1] * cl2[:Lmax + 1]
        return ret

    def _get_n1_L(self, L, kA, kB, k_ind, cl_kind, ftlA, felA, fblA, ftlB, felB, fblB, clttfid, cltefid, cleefid,
                  remove_only=False):
        """N1 at one multipole, from the float cache (reference: n1.py:273-316)."""
        if kB is None:
            kB = kA
        assert kA in estimator_keys and kB in estimator_keys
        assert len(cl_kind) > self.lmaxphi
        if kA < kB:
            return self._get_n1_L(L, kB, kA, k_ind, cl_kind, ftlB, felB, fblB, ftlA, felA, fblA, clttfid, cltefid, cleefid)
        lmax_ftl = np.max([len(fal) for fal in [ftlA, felA, fblA, ftlB, felB, fblB]]) - 1
        for cl_fid, cl_map in ((clttfid, self.cltt), (cltefid, self.clte), (cleefid, self.clee)):
            assert len(cl_fid) > lmax_ftl and len(cl_map) > lmax_ftl
        idx = self._key(str(L), kA, kB, k_ind, cl_kind, (ftlA, felA, fblA, ftlB, felB, fblB), (clttfid, cltefid, cleefid))
        n1_L = self.fldb.get(idx)
        if n1_L is None:
            if remove_only:
                return 0.
            if self.n1l is None:
                raise NotImplementedError("N1 at L = %s for (%s, %s) is not in %s and this package has no flat-sky N1 "
                                          "integrator (the reference's n1f Fortran extension)" % (L, kA, kB, self.lib_dir))
            lmin_A = np.min([np.where(np.abs(fal) > 0.)[0][0] for fal in [ftlA, felA, fblA]])
            lmin_B = np.min([np.where(np.abs(fal) > 0.)[0][0] for fal in [ftlB, felB, fblB]])
            n1_L = self.n1l(L, cl_kind, kA, kB, k_ind, self.cltt, self.clte, self.clee, clttfid, cltefid, cleefid,
                            ftlA, felA, fblA, ftlB, felB, fblB, lmin_A, lmin_B, self.dL, self.lps)
            self.fldb.add(idx, n1_L)
            return n1_L
        if remove_only:
            self.fldb.remove(idx)
            return 0.
        return n1_L

    def get_n1_jtp(self, *args, **kwargs):
        """Joint T-P filtering variant (reference: n1.py:318-390): needs the integrator for every term."""
        raise NotImplementedError("get_n1_jtp needs the flat-sky N1 integrator (the reference's n1f Fortran extension)")
