"""Semi-analytical Gaussian noise levels (N0) of the quadratic estimators (reference: plancklens/nhl.py:15-96).

`get_nhl` is a sum of Wigner-d correlation functions of the estimator weights and the filtered-map spectra; the
transforms run on the GPU (`utils_spin.wignerc` -> libplk_b200) where the reference calls its Fortran extension.
`nhl_lib_simple` evaluates it on the spectra of the filtered simulations and caches the result in the reference's
sqlite layout (`helpers/sql.py`).
"""
import os
import pickle as pk

import numpy as np

from . import hp, qresp, utils
from . import utils_spin as uspin
from .helpers import mpi, sql


def get_nhl(qe_key1, qe_key2, cls_weights, cls_ivfs, lmax_ivf1, lmax_ivf2, lmax_out=None, lmax_ivf12=None,
            lmax_ivf22=None, cls_weights2=None, cls_ivfs_bb=None, cls_ivfs_ab=None, cls_ivfs_ba=None):
    """(Semi-)analytical noise level of the cross-spectrum of two QE keys: (GG, CC, GC, CG) (reference: nhl.py:15-43)."""
    if lmax_ivf12 is None:
        lmax_ivf12 = lmax_ivf1
    if lmax_ivf22 is None:
        lmax_ivf22 = lmax_ivf2
    if cls_weights2 is None:
        cls_weights2 = cls_weights
    qes1 = qresp.get_qes(qe_key1, lmax_ivf1, cls_weights, lmax2=lmax_ivf12)
    qes2 = qresp.get_qes(qe_key2, lmax_ivf2, cls_weights2, lmax2=lmax_ivf22)
    if lmax_out is None:
        lmax_out = max(lmax_ivf1, lmax_ivf12) + max(lmax_ivf2, lmax_ivf22)
    return _get_nhl(qes1, qes2, cls_ivfs, lmax_out, cls_ivfs_bb=cls_ivfs_bb, cls_ivfs_ab=cls_ivfs_ab, cls_ivfs_ba=cls_ivfs_ba)


def _get_nhl(qes1, qes2, cls_ivfs, lmax_out, cls_ivfs_bb=None, cls_ivfs_ab=None, cls_ivfs_ba=None):
    """reference: nhl.py:45-96 (same loops, same order of accumulation)"""
    GG_N0 = np.zeros(lmax_out + 1, dtype=float)
    CC_N0 = np.zeros(lmax_out + 1, dtype=float)
    GC_N0 = np.zeros(lmax_out + 1, dtype=float)
    CG_N0 = np.zeros(lmax_out + 1, dtype=float)
    cls_ivfs_aa = cls_ivfs
    cls_ivfs_bb = cls_ivfs if cls_ivfs_bb is None else cls_ivfs_bb
    cls_ivfs_ab = cls_ivfs if cls_ivfs_ab is None else cls_ivfs_ab
    cls_ivfs_ba = cls_ivfs if cls_ivfs_ba is None else cls_ivfs_ba
    j, w, sc = utils.joincls, uspin.wignerc, uspin.spin_cls
    Ls = np.arange(lmax_out + 1)
    for qe1 in qes1:
        cL1 = qe1.cL(Ls)
        for qe2 in qes2:
            cL2 = qe2.cL(Ls)
            si, ti, ui, vi = (qe1.leg_a.spin_in, qe1.leg_b.spin_in, qe2.leg_a.spin_in, qe2.leg_b.spin_in)
            so, to, uo, vo = (qe1.leg_a.spin_ou, qe1.leg_b.spin_ou, qe2.leg_a.spin_ou, qe2.leg_b.spin_ou)
            assert so + to >= 0 and uo + vo >= 0, (so, to, uo, vo)
            a1, b1, a2, b2 = qe1.leg_a.cl, qe1.leg_b.cl, qe2.leg_a.cl, qe2.leg_b.cl

            clsu = j([a1, a2.conj(), sc(si, ui, cls_ivfs_aa)])
            cltv = j([b1, b2.conj(), sc(ti, vi, cls_ivfs_bb)])
            R_sutv = j([w(clsu, cltv, so, uo, to, vo, lmax_out=lmax_out), cL1, cL2])
            clsv = j([a1, b2.conj(), sc(si, vi, cls_ivfs_ab)])
            cltu = j([b1, a2.conj(), sc(ti, ui, cls_ivfs_ba)])
            R_sutv = R_sutv + j([w(clsv, cltu, so, vo, to, uo, lmax_out=lmax_out), cL1, cL2])

            # -s -t u v
            sgnms = (-1) ** (si + so)
            sgnmt = (-1) ** (ti + to)
            clsu = j([sgnms * a1.conj(), a2.conj(), sc(-si, ui, cls_ivfs_aa)])
            cltv = j([sgnmt * b1.conj(), b2.conj(), sc(-ti, vi, cls_ivfs_bb)])
            R_msmtuv = j([w(clsu, cltv, -so, uo, -to, vo, lmax_out=lmax_out), cL1, cL2])
            clsv = j([sgnms * a1.conj(), b2.conj(), sc(-si, vi, cls_ivfs_ab)])
            cltu = j([sgnmt * b1.conj(), a2.conj(), sc(-ti, ui, cls_ivfs_ba)])
            R_msmtuv = R_msmtuv + j([w(clsv, cltu, -so, vo, -to, uo, lmax_out=lmax_out), cL1, cL2])

            sg = (-1) ** (to + so)
            GG_N0 += 0.5 * R_sutv.real + 0.5 * sg * R_msmtuv.real
            CC_N0 += 0.5 * R_sutv.real - 0.5 * sg * R_msmtuv.real
            GC_N0 -= 0.5 * R_sutv.imag + 0.5 * sg * R_msmtuv.imag
            CG_N0 += 0.5 * R_sutv.imag - 0.5 * sg * R_msmtuv.imag
    return GG_N0, CC_N0, GC_N0, CG_N0


class nhl_lib_simple:
    """Semi-analytical unnormalised N0 library for four identical legs and the 1 / fsky spectrum estimator
    (reference: nhl.py:98-189)."""

    def __init__(self, lib_dir, ivfs, cls_weight, lmax_qlm, resplib=None):
        self.lmax_qlm = lmax_qlm
        self.cls_weight = cls_weight
        self.ivfs = ivfs
        fn_hash = os.path.join(lib_dir, 'nhl_hash.pk')
        if mpi.rank == 0:
            if not os.path.exists(lib_dir):
                os.makedirs(lib_dir)
            if not os.path.exists(fn_hash):
                with open(fn_hash, 'wb') as f:
                    pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fn_hash, 'rb') as f:
            utils.hash_check(pk.load(f), self.hashdict(), fn=fn_hash)
        self.lib_dir = lib_dir
        self.npdb = sql.npdb(os.path.join(lib_dir, 'npdb.db'))
        self.fsky = np.mean(self.ivfs.get_fmask())
        self.resplib = resplib

    def hashdict(self):
        ret = {k: utils.clhash(self.cls_weight[k]) for k in self.cls_weight.keys()}
        ret['ivfs'] = self.ivfs.hashdict()
        ret['lmax_qlm'] = self.lmax_qlm
        return ret

    def _get_qe_derived(self, k):
        if '_bh_' in k:
            kQE, ksource = k.split('_bh_')
            assert len(ksource) == 1
            wL = self.resplib.get_response(kQE, ksource) * utils.cli(self.resplib.get_response(ksource + kQE[1:], ksource))
            return [(kQE, 1.), (ksource + kQE[1:], -wL)]
        return [(k, 1.)]

    def get_sim_nhl(self, idx, k1, k2, recache=False):
        """N0 of the (k1, k2) spectrum from the empirical spectra of the filtered maps of simulation idx (-1: data)."""
        assert idx == -1 or idx >= 0, idx
        ret = np.zeros(self.lmax_qlm + 1)
        for k1_, w1 in self._get_qe_derived(k1):
            for k2_, w2 in self._get_qe_derived(k2):
                s1, GC1, s1ins, ksp1 = qresp.qe_spin_data(k1_)
                s2, GC2, s2ins, ksp2 = qresp.qe_spin_data(k2_)
                base = 'anhl_qe_' + ksp1 + k1_[1:] + '_qe_' + ksp2 + k2_[1:]
                suf = ('sim%04d' % idx) * (int(idx) >= 0) + 'dat' * (idx == -1)
                fn = base + GC1 + GC2
                if self.npdb.get(fn + suf) is None or recache:
                    assert s1 >= 0 and s2 >= 0, (s1, s2)
                    cls_ivfs, lmax_ivf = self._get_cls(idx, np.unique(np.concatenate([s1ins, s2ins])))
                    GG, CC, GC, CG = get_nhl(k1_, k2_, self.cls_weight, cls_ivfs, lmax_ivf, lmax_ivf, lmax_out=self.lmax_qlm)
                    fns = [('G', 'G', GG)] + [('C', 'G', CG)] * (s1 > 0) + [('G', 'C', GC)] * (s2 > 0) \
                        + [('C', 'C', CC)] * (s1 > 0) * (s2 > 0)
                    if recache and self.npdb.get(fn + suf) is not None:
                        for a, b, _ in fns:
                            self.npdb.remove(base + a + b + suf)
                    for a, b, N0 in fns:
                        self.npdb.add(base + a + b + suf, N0)
                ret += w1 * w2 * self.npdb.get(fn + suf)
        return ret

    def _get_cls(self, idx, spins):
        """empirical spectra of the filtered alms / fsky, and the length the reference passes on as lmax (nhl.py:175-189)"""
        assert np.all(spins >= 0), spins
        ret = {}
        iv = self.ivfs
        if 0 in spins:
            ret['tt'] = hp.alm2cl(iv.get_sim_tlm(idx)) / self.fsky
        if 2 in spins:
            ret['ee'] = hp.alm2cl(iv.get_sim_elm(idx)) / self.fsky
            ret['bb'] = hp.alm2cl(iv.get_sim_blm(idx)) / self.fsky
            ret['eb'] = hp.alm2cl(iv.get_sim_elm(idx), alms2=iv.get_sim_blm(idx)) / self.fsky
        if 0 in spins and 2 in spins:
            ret['te'] = hp.alm2cl(iv.get_sim_tlm(idx), alms2=iv.get_sim_elm(idx)) / self.fsky
            ret['tb'] = hp.alm2cl(iv.get_sim_tlm(idx), alms2=iv.get_sim_blm(idx)) / self.fsky
        lmaxs = [len(cl) for cl in ret.values()]
        assert len(np.unique(lmaxs)) == 1, lmaxs
        return ret, lmaxs[0]


def cls2dls(cls):
    """see n0s.cls2dls (the reference keeps a second copy here, nhl.py:191-203)"""
    from . import n0s
    return n0s.cls2dls(cls)


def dls2cls(dls):
    from . import n0s
    return n0s.dls2cls(dls)


def get_N0_iter(*args, **kwargs):
    from . import n0s
    return n0s.get_N0_iter(*args, **kwargs)
