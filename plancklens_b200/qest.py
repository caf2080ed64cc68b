"""Quadratic-estimator evaluation (reference: plancklens/qest.py), GPU resident.

`library.get_sim_qlm(k, idx)` keeps the reference's behaviour (keys, caching under lib_dir, symmetrisation when
the two legs come from different filtering libraries).  For the lensing keys 'ptt'/'xtt', 'p_p'/'x_p', 'p'/'x'
the legs are synthesised, multiplied pixel by pixel and analysed back entirely on the GPU:

  T estimator (qest.py:248-263):   alm2map(T_bar)                       ->  t
                                   alm2map_spin(-sqrt(l(l+1)) T^WF, 1)  ->  (G, C);   (G t, C t)
  P estimator (qest.py:265-285):   alm2map_spin(1/2 E_bar, 1/2 B_bar, 2) ->  (Q, U)
                                   alm2map_spin(sqrt((l-2)(l+3)) (E,B)^WF, 3) -> (G3, C3)
                                   alm2map_spin(sqrt((l+2)(l-1)) (E,B)^WF, 1) -> (G1, C1)
                                   (Q - iU)(G3 + iC3) - (Q + iU)(G1 - iC1)
  then map2alm_spin(., 1, lmax_qlm) x -sqrt(L(L+1)).

All per-l factors are fused into the transforms (`fl` arguments of libplk_b200).  For the MV key 'p' the T and P
real-space products are summed BEFORE one spin-1 analysis (the analysis is linear, so this equals the reference's
sum of two analyses to round-off and saves one of seven transforms); `merge_analysis=False` restores the
reference's order of operations.
"""
import collections
import os
import pickle as pk

import numpy as np
import torch

from . import hp, sht
from . import utils as ut
from .helpers import mpi

_write_alm = lambda fn, alm: hp.write_alm(fn, alm, overwrite=True)
_XY_PAIRS = ['te', 'et', 'tb', 'bt', 'ee', 'eb', 'be', 'bb']            # single-pair lensing keys 'pte', 'xeb', ...
_SYM_PAIR_KEYS = ['p_te', 'p_tb', 'p_eb', 'x_te', 'x_tb', 'x_eb']


def eval_qe(qe_key, lmax_ivf, cls_weight, get_alm, nside, lmax_qlm, verbose=True, get_alm2=None, transf=None):
    """Gradient and curl terms of a quadratic estimator through the generic leg machinery
    (reference: qest.py:19-39; `library` below is faster for the lensing keys).

        Args:
            qe_key: estimator key as defined in `qresp` (e.g. 'ptt')
            lmax_ivf: CMB multipoles up to lmax_ivf are used
            cls_weight: CMB spectra entering the estimator weights
            get_alm: callable with 't', 'e', 'b' returning the inverse-variance filtered alms
            nside: resolution of the real-space products
            lmax_qlm: maximum multipole of the estimate
            get_alm2: alms of the second leg if different (the estimator is then symmetrised)
    """
    from . import qresp, utils_qe
    qe_list = qresp.get_qes(qe_key, lmax_ivf, cls_weight, transf=transf)
    return utils_qe.qe_eval(qe_list, nside, get_alm, lmax_qlm, verbose=verbose, get_alm2=get_alm2)


def library_jtTP(lib_dir, ivfs1, ivfs2, nside, lmax_qlm=None, resplib=None):
    return library(lib_dir, ivfs1, ivfs2, nside, lmax_qlm=lmax_qlm, resplib=resplib)


def library_sepTP(lib_dir, ivfs1, ivfs2, clte, nside, lmax_qlm=None, resplib=None):
    return library(lib_dir, ivfs1, ivfs2, nside, clte=clte, lmax_qlm=lmax_qlm, resplib=resplib)


def _dfl(fl):
    return torch.from_numpy(np.ascontiguousarray(fl, dtype=np.float64)).cuda()


def _combine(lmax, terms):
    """sum_j fl_j[l] a_j[l,m] on the device (plk_alm_combine_dev)."""
    import ctypes
    n = len(terms)
    out = torch.empty_like(terms[0][0])
    ins = (ctypes.c_void_p * n)(*[t[0].data_ptr() for t in terms])
    fls = (ctypes.c_void_p * n)(*[t[1].data_ptr() for t in terms])
    nfl = (ctypes.c_int * n)(*[int(t[1].numel()) for t in terms])
    sht.check(sht._lib.load().plk_alm_combine_dev(lmax, n, ins, fls, nfl, sht._ptr(out), sht._stream()))
    return out


def _grad_fl(lmax, spin, kind):
    """per-l factors of the gradient legs (reference: qest.py:463, 494-503)."""
    l = np.arange(lmax + 1, dtype=float)
    if kind == 't':
        return -np.sqrt(l * (l + 1))
    fl = (l + 2) * (l - 1) if spin == 1 else (l - 2) * (l + 3)
    fl[:spin] = 0.
    return np.sqrt(np.maximum(fl, 0.))


class qe_device:
    """The GPU evaluation of the lensing estimators from device-resident filtered alms.

    tbar, ebar, bbar: inverse-variance filtered alms; twf, ewf, bwf: Wiener-filtered alms entering the gradient
    legs (already including the C^TE cross terms when relevant).  All are complex128 CUDA tensors of lmax_ivf.
    """

    def __init__(self, nside, lmax_ivf, lmax_qlm, plan_ivf=None, plan_qlm=None):
        """plan_ivf / plan_qlm: objects with the `sht.Plan` transform methods; pass `dist_sht.DistPlan`s to split every
        transform of one estimate over the GPUs of a box by m (BASELINE.json configs[4])."""
        self.nside, self.lmax_ivf, self.lmax_qlm = nside, lmax_ivf, lmax_qlm
        self.plan_ivf = sht.get_plan(nside, lmax_ivf) if plan_ivf is None else plan_ivf
        self.plan_qlm = sht.get_plan(nside, lmax_qlm) if plan_qlm is None else plan_qlm
        self.fl_t1 = _dfl(_grad_fl(lmax_ivf, 1, 't'))
        self.fl_p1 = _dfl(_grad_fl(lmax_ivf, 1, 'p'))
        self.fl_p3 = _dfl(_grad_fl(lmax_ivf, 3, 'p'))
        self.fl_half = _dfl(0.5 * np.ones(lmax_ivf + 1))
        L = np.arange(lmax_qlm + 1, dtype=float)
        self.fl_out = _dfl(-np.sqrt(L * (L + 1)))
        npix = 12 * nside ** 2
        self._buf = [torch.empty(npix, dtype=torch.float64, device='cuda') for _ in range(10)]
        # m-partitioned plans: this process owns (and multiplies) only the pixels of its own rings
        pr = getattr(self.plan_ivf, 'pixel_ranges', None)
        self._pix = [slice(lo, hi) for lo, hi in pr() if hi > lo] if (pr is not None and hasattr(self.plan_ivf, 'rank')) \
            else [slice(None)]

    def t_products(self, tbar, twf, out=None):
        """(G t, C t) maps of the temperature estimator."""
        b = self._buf
        t = self.plan_ivf.alm2map(tbar, out=b[0])
        G, C = self.plan_ivf.alm2map_spin(twf, None, 1, flg=self.fl_t1, out=(b[1], b[2]) if out is None else out)
        for s in self._pix:
            sht.map_mul2(G[s], C[s], t[s])
        return G, C

    def p_products(self, ebar, bbar, ewf, bwf, out=None):
        """(Re, Im) maps of the polarization estimator."""
        b = self._buf
        Q, U = self.plan_ivf.alm2map_spin(ebar, bbar, 2, flg=self.fl_half, flc=self.fl_half, out=(b[0], b[3]))
        G3, C3 = self.plan_ivf.alm2map_spin(ewf, bwf, 3, flg=self.fl_p3, flc=self.fl_p3, out=(b[4], b[5]))
        G1, C1 = self.plan_ivf.alm2map_spin(ewf, bwf, 1, flg=self.fl_p1, flc=self.fl_p1, out=(b[6], b[7]))
        re, im = (b[8], b[9]) if out is None else out
        for s in self._pix:
            sht.map_qe_pp(Q[s], U[s], G3[s], C3[s], G1[s], C1[s], re[s], im[s])
        return re, im

    def analyse(self, re, im):
        return self.plan_qlm.map2alm_spin(re, im, 1, flg=self.fl_out, flc=self.fl_out)

    def to_host(self, G, C):
        """device qlm pair -> numpy, through pinned staging buffers"""
        if not hasattr(self, '_pin'):
            self._pin = [torch.empty(G.numel(), dtype=torch.complex128, pin_memory=True) for _ in range(2)]
        self._pin[0].copy_(G, non_blocking=True)
        self._pin[1].copy_(C, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pin[0].numpy().copy(), self._pin[1].numpy().copy()

    # ---- leg products evaluated INSIDE the analysis ring kernel (plk_map2alm_pix_dev): the product maps of
    #      qest.py:256-257 and :276-278 are never written; single-GPU plans only (the m-partitioned ring stage keeps the
    #      separate product kernels)
    def _fused(self):
        return hasattr(self.plan_qlm, 'map2alm_spin_pix') and os.environ.get('PLK_QE_FUSED', '1') != '0'

    def _t_legs(self, tbar, twf):
        b = self._buf
        t = self.plan_ivf.alm2map(tbar, out=b[0])
        G, C = self.plan_ivf.alm2map_spin(twf, None, 1, flg=self.fl_t1, out=(b[1], b[2]))
        return [(1.0, G, t)], [(1.0, C, t)]

    def _p_legs(self, ebar, bbar, ewf, bwf):
        b = self._buf
        Q, U = self.plan_ivf.alm2map_spin(ebar, bbar, 2, flg=self.fl_half, flc=self.fl_half, out=(b[3], b[8]))
        G3, C3 = self.plan_ivf.alm2map_spin(ewf, bwf, 3, flg=self.fl_p3, flc=self.fl_p3, out=(b[4], b[5]))
        G1, C1 = self.plan_ivf.alm2map_spin(ewf, bwf, 1, flg=self.fl_p1, flc=self.fl_p1, out=(b[6], b[7]))
        # (Q - iU)(G3 + iC3) - (Q + iU)(G1 - iC1)
        re = [(1.0, Q, G3), (1.0, U, C3), (-1.0, Q, G1), (-1.0, U, C1)]
        im = [(1.0, Q, C3), (-1.0, U, G3), (-1.0, U, G1), (1.0, Q, C1)]
        return re, im

    def _analyse_fused(self, re, im):
        return self.plan_qlm.map2alm_spin_pix(re, im, 1, flg=self.fl_out, flc=self.fl_out)

    def ptt(self, tbar, twf):
        if self._fused():
            return self._analyse_fused(*self._t_legs(tbar, twf))
        return self.analyse(*self.t_products(tbar, twf))

    def p_p(self, ebar, bbar, ewf, bwf):
        if self._fused():
            return self._analyse_fused(*self._p_legs(ebar, bbar, ewf, bwf))
        return self.analyse(*self.p_products(ebar, bbar, ewf, bwf))

    def p(self, tbar, ebar, bbar, twf, ewf, bwf, merge_analysis=True):
        if merge_analysis and self._fused():
            rp, ip = self._p_legs(ebar, bbar, ewf, bwf)
            rt, it = self._t_legs(tbar, twf)
            return self._analyse_fused(rp + rt, ip + it)        # 5 product terms per component, one spin-1 analysis
        b = self._buf
        re, im = self.p_products(ebar, bbar, ewf, bwf)             # in b[8], b[9]
        if merge_analysis:
            G, C = self.t_products(tbar, twf)                       # in b[1], b[2]
            for s in self._pix:
                sht.map_axpy(re[s], G[s], 1.0)
                sht.map_axpy(im[s], C[s], 1.0)
            return self.analyse(re, im)
        GP, CP = self.analyse(re, im)
        GT, CT = self.analyse(*self.t_products(tbar, twf))
        sht.alm_axpy(GP, GT, 1.0)
        sht.alm_axpy(CP, CT, 1.0)
        return GP, CP


class library:
    r"""QE evaluation library built on two inverse-variance filtered simulation libraries
    (reference: qest.py:50-438).

        Args:
            lib_dir: estimates are cached there
            ivfs1, ivfs2: filtering libraries of the first and second leg
            nside: resolution of the real-space products
            clte (optional): TE spectrum used to build the Wiener-filtered legs from separately filtered T and P
            lmax_qlm (optional): maximum multipole of the estimates (default 3 nside - 1)
    """

    def __init__(self, lib_dir, ivfs1, ivfs2, nside, clte=None, lmax_qlm=None, resplib=None, merge_analysis=True):
        if lmax_qlm is None:
            lmax_qlm = 3 * nside - 1
        self.lib_dir = lib_dir
        self.prefix = lib_dir
        self.nside = nside
        self.lmax_qlm = {'T': lmax_qlm, 'P': lmax_qlm, 'PS': lmax_qlm}
        self.merge_analysis = merge_analysis
        if clte is None:
            self.f2map1 = lib_filt2map(ivfs1, nside)
            self.f2map2 = lib_filt2map(ivfs2, nside)
        else:
            self.f2map1 = lib_filt2map_sepTP(ivfs1, nside, clte)
            self.f2map2 = lib_filt2map_sepTP(ivfs2, nside, clte)
        fnhash = os.path.join(self.lib_dir, "qe_sim_hash.pk")
        if mpi.rank == 0 and not os.path.exists(fnhash):
            if not os.path.exists(self.lib_dir):
                os.makedirs(self.lib_dir)
            with open(fnhash, 'wb') as f:
                pk.dump(self.hashdict(), f, protocol=2)
        mpi.barrier()
        with open(fnhash, 'rb') as f:
            ut.hash_check(pk.load(f), self.hashdict(), fn=fnhash)
        fn_fsky = os.path.join(lib_dir, 'fskies.dat')
        if mpi.rank == 0 and not os.path.exists(fn_fsky):
            ms = {1: self.get_mask(1), 2: self.get_mask(2)}
            with open(fn_fsky, 'w') as f:
                for i in [1, 2]:
                    for j in [1, 2][i - 1:]:
                        f.write('%4s %.5f \n' % (10 * i + j, np.mean(ms[i] * ms[j])))
        mpi.barrier()
        self.fskies = {}
        with open(fn_fsky) as f:
            for line in f:
                key, val = line.split()
                self.fskies[int(key)] = float(val)
        self.fsky11, self.fsky12, self.fsky22 = self.fskies[11], self.fskies[12], self.fskies[22]
        self.resplib = resplib
        self.keys_fund = ['ptt', 'xtt', 'p_p', 'x_p', 'p', 'x', 'stt', 's', 'ftt', 'f_p', 'f', 'dtt', 'ntt', 'a_p'] \
                         + [s + xy for s in 'px' for xy in _XY_PAIRS]
        self.keys = self.keys_fund + ['p_tp', 'x_tp', 'p_te', 'p_tb', 'p_eb', 'x_te', 'x_tb', 'x_eb', 'ptt_bh_n',
                                      'ptt_bh_s', 'ptt_bh_f', 'ptt_bh_d', 'dtt_bh_p', 'stt_bh_p', 'ftt_bh_d', 'p_bh_s']
        self.keys_remaps = {'s': 'stt'}      # equivalent keys (reference: qest.py:118)
        self._qe = None

    def hashdict(self):
        return {'f2map1': self.f2map1.hashdict(), 'f2map2': self.f2map2.hashdict()}

    def get_fundkeys(self, k_list):
        klist = k_list if isinstance(k_list, list) else [k_list]
        ret = []
        for k in klist:
            if k in self.keys_fund:
                ret.append(k)
            elif '_tp' in k:
                ret += [k[0] + 'tt', k[0] + '_p']
            elif 'tt_bh_' in k:
                kqe, src = k.split('_bh_')
                ret += [kqe, src + 'tt']
            elif k in _SYM_PAIR_KEYS:
                ret += [k[0] + k[2] + k[3], k[0] + k[3] + k[2]]
        return list(collections.OrderedDict.fromkeys(ret))

    def get_fsky(self, id):
        assert id in [11, 22, 12], id
        return self.fskies[id]

    def get_lmax_qlm(self, k):
        assert self.lmax_qlm['T'] == self.lmax_qlm['P']
        return self.lmax_qlm['T']

    def get_mask(self, leg):
        assert leg in [1, 2]
        return self.f2map1.ivfs.get_fmask() if leg == 1 else self.f2map2.ivfs.get_fmask()

    def get_sim_qlm(self, k, idx, lmax=None):
        """QE estimate for key k and simulation idx (computed on the GPU and cached on first call)."""
        k = self.keys_remaps.get(k, k)
        if lmax is None:
            lmax = self.get_lmax_qlm(k)
        assert lmax <= self.get_lmax_qlm(k)
        if k in ['p_tp', 'x_tp', 'f_tp', 's_tp']:
            return self.get_sim_qlm('%stt' % k[0], idx, lmax=lmax) + self.get_sim_qlm('%s_p' % k[0], idx, lmax=lmax)
        if k in _SYM_PAIR_KEYS:             # e.g. 'p_eb' = 'peb' + 'pbe'
            return self.get_sim_qlm(k[0] + k[2] + k[3], idx, lmax=lmax) + self.get_sim_qlm(k[0] + k[3] + k[2], idx, lmax=lmax)
        if '_bh_' in k:
            kQE, kS, wL = self._bh_split(k)
            lmax = self.get_lmax_qlm(kQE)
            return self.get_sim_qlm(kQE, idx, lmax=lmax) - hp.almxfl(self.get_sim_qlm(kS, idx, lmax=lmax), wL)
        assert k in self.keys_fund, (k, self.keys_fund)
        fname = os.path.join(self.lib_dir, 'sim_%s_%04d.fits' % (k, idx) if idx != -1 else 'dat_%s.fits' % k)
        if not os.path.exists(fname):
            if k in ['ptt', 'xtt']:
                self._build_sim_Tgclm(idx)
            elif k in ['p_p', 'x_p']:
                self._build_sim_Pgclm(idx)
            elif k in ['p', 'x']:
                self._build_sim_MVgclm(idx)
            elif k in ['f', 'stt', 'ftt', 'f_p', 'ntt', 'a_p']:
                getattr(self, '_build_sim_' + k)(idx)
            elif k[1:] in _XY_PAIRS:
                self._build_sim_xfiltMVgclm(idx, k)
            else:
                assert 0, k
        return ut.alm_copy(hp.read_alm(fname), lmax=lmax)

    def get_dat_qlm(self, k, **kwargs):
        return self.get_sim_qlm(k, -1, **kwargs)

    def _bh_split(self, k):
        """'ptt_bh_s' -> ('ptt', 'stt', w_L): the estimator hardened against source s is
        qlm(ptt) - w_L qlm(stt) with w_L = R^{ptt,s}_L / R^{stt,s}_L from the response library (reference: qest.py:170-177)."""
        assert self.resplib is not None, 'resplib arg necessary for this'
        kQE, src = k.split('_bh_')
        assert len(src) == 1, (src, kQE)
        kS = src + kQE[1:]
        assert self.get_lmax_qlm(kQE) == self.get_lmax_qlm(kS), (kQE, kS)
        return kQE, kS, self.resplib.get_response(kQE, src) * ut.cli(self.resplib.get_response(kS, src))

    def get_sim_qlm_mf(self, k, mc_sims, lmax=None):
        """Mean-field estimate: average of the QE over mc_sims (reference: qest.py:206-246)."""
        k = self.keys_remaps.get(k, k)
        if lmax is None:
            lmax = self.get_lmax_qlm(k)
        assert lmax <= self.get_lmax_qlm(k)
        if k in ['p_tp', 'x_tp']:
            return self.get_sim_qlm_mf('%stt' % k[0], mc_sims, lmax=lmax) + self.get_sim_qlm_mf('%s_p' % k[0], mc_sims, lmax=lmax)
        if k in _SYM_PAIR_KEYS:
            return self.get_sim_qlm_mf(k[0] + k[2] + k[3], mc_sims, lmax=lmax) + self.get_sim_qlm_mf(k[0] + k[3] + k[2], mc_sims, lmax=lmax)
        if '_bh_' in k:
            kQE, kS, wL = self._bh_split(k)
            lmax = self.get_lmax_qlm(kQE)
            return self.get_sim_qlm_mf(kQE, mc_sims, lmax=lmax) - hp.almxfl(self.get_sim_qlm_mf(kS, mc_sims, lmax=lmax), wL)
        assert k in self.keys_fund, (k, self.keys_fund)
        fname = os.path.join(self.lib_dir, 'simMF_k1%s_%s.fits' % (k, ut.mchash(mc_sims)))
        if not os.path.exists(fname):
            # a purely LOCAL call, as in the reference: the calling rank averages all of mc_sims itself.  Drivers hand
            # different (k, mc_sims) jobs to different ranks (examples/run_qlms.py:92-94) and qecl.get_sim_qcl calls
            # this from rank-strided loops (qecl.py:85-86), so a collective hidden here would pair up unrelated
            # reductions or hang.  The sharded form is `get_sim_qlm_mf_sharded`, which says so in its name.
            this_mcs = np.unique(mc_sims)
            MF = np.zeros(hp.Alm.getsize(lmax), dtype=complex)
            if len(this_mcs) == 0:
                return MF
            for idx in this_mcs:
                MF += self.get_sim_qlm(k, idx, lmax=lmax)
            MF /= len(this_mcs)
            _write_alm(fname, MF)
        return ut.alm_copy(hp.read_alm(fname), lmax=lmax)

    def get_sim_qlm_mf_sharded(self, k, mc_sims, lmax=None):
        """COLLECTIVE mean field of a fundamental key: every rank of the process group must call it with the same
        arguments.  Simulations are strided over ranks (`this_mcs[rank::size]`, the pattern of
        examples/run_qlms.py:72), the partial sums are reduced over NCCL / gloo -- the serial loop of qest.py:239-243
        made parallel -- rank 0 writes the reference's cache file and all ranks return the
        same array.  Not a drop-in for `get_sim_qlm_mf`, which stays local."""
        k = self.keys_remaps.get(k, k)
        if lmax is None:
            lmax = self.get_lmax_qlm(k)
        assert lmax <= self.get_lmax_qlm(k)
        assert k in self.keys_fund, (k, self.keys_fund)
        mpi.require_group()      # raises when WORLD_SIZE > 1 but no process group can be joined
        fname = os.path.join(self.lib_dir, 'simMF_k1%s_%s.fits' % (k, ut.mchash(mc_sims)))
        have = mpi.bcast(os.path.exists(fname) if mpi.rank == 0 else None, root=0)
        if not have:
            this_mcs = np.unique(mc_sims)
            MF = np.zeros(hp.Alm.getsize(lmax), dtype=complex)
            if len(this_mcs) == 0:
                return MF
            for idx in this_mcs[mpi.rank::mpi.size]:
                MF += self.get_sim_qlm(k, idx, lmax=lmax)
            MF = mpi.allreduce_sum(MF) / len(this_mcs)
            if mpi.rank == 0:
                _write_alm(fname, MF)
            mpi.barrier()
        return ut.alm_copy(hp.read_alm(fname), lmax=lmax)

    def eval_qlm(self, k, idx, swapped=False):
        """Gradient and curl estimate for a fundamental key, computed on the GPU, not cached (numpy out)."""
        assert k in ['ptt', 'p_p', 'p'], k
        fun = {'ptt': self._get_sim_Tgclm, 'p_p': self._get_sim_Pgclm, 'p': self._get_sim_MVgclm}[k]
        return fun(idx, k, swapped=swapped)

    def eval_qlms(self, k, idxs):
        """Pipelined `eval_qlm` over many simulations: yields (idx, G, C) in order (numpy out, same legs on both
        sides -- the `qlms_dd` case).

        Three things run concurrently: the host -> device copy of simulation i + 1's filtered alms (upload stream),
        the transforms of simulation i (compute stream), and the device -> host copy (download stream) plus the caller's handling of
        estimate i - 1.  Inputs and outputs are double buffered; pinned host arrays (torch `pin_memory`) are copied
        asynchronously, pageable ones go through pinned staging buffers."""
        assert k in ['ptt', 'p_p', 'p'], k
        assert self.f2map1.ivfs is self.f2map2.ivfs, "pipelined evaluation covers identical legs; use eval_qlm otherwise"
        ivfs, f2 = self.f2map1.ivfs, self.f2map2
        # one stream per copy direction: on a single copy stream the upload of simulation i + 1 would queue behind the
        # download of estimate i, which waits for the transforms of simulation i -- no overlap left
        main, cs, cs_out = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
        h2d_done, comp_done, d2h_done = ev(), ev(), ev()
        dev_in, pin_in, pin_out, live = [None, None], [None, None], [None, None], [None, None]
        names = {'ptt': ('t',), 'p_p': ('e', 'b'), 'p': ('t', 'e', 'b')}[k]

        def fetch(idx):
            return [getattr(ivfs, 'get_sim_%slm' % n)(idx) for n in names]

        def finish(slot, idx):
            d2h_done[slot].synchronize()
            G, C = pin_out[slot][0].numpy().copy(), pin_out[slot][1].numpy().copy()
            live[slot] = None
            return idx, G, C

        prev = None
        for i, idx in enumerate(idxs):
            s = i % 2
            host = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)) for a in fetch(idx)]
            if dev_in[s] is None:
                dev_in[s] = [torch.empty(h.numel(), dtype=torch.complex128, device='cuda') for h in host]
                pin_in[s] = [None] * len(host)
            if i >= 2:
                cs.wait_event(comp_done[s])              # the transforms of simulation i - 2 have consumed this slot
            with torch.cuda.stream(cs):
                for j, h in enumerate(host):
                    if not h.is_pinned():
                        if pin_in[s][j] is None:
                            pin_in[s][j] = torch.empty(h.numel(), dtype=torch.complex128, pin_memory=True)
                        if i >= 2:
                            h2d_done[s].synchronize()    # staging buffer free again
                        pin_in[s][j].copy_(h)
                        h = pin_in[s][j]
                    dev_in[s][j].copy_(h, non_blocking=True)
                h2d_done[s].record(cs)
            live[s] = [host]                             # keep the host arrays alive until the copy has run
            main.wait_event(h2d_done[s])
            qe = self._engine(sht.alm_lmax(dev_in[s][0].numel()))
            if k == 'ptt':
                (dt,) = dev_in[s]
                G, C = qe.ptt(dt, sht.almxfl(dt, f2._cl_dev()['tt']))
            elif k == 'p_p':
                de, db = dev_in[s]
                c = f2._cl_dev()
                G, C = qe.p_p(de, db, sht.almxfl(de, c['ee']), sht.almxfl(db, c['bb']))
            else:
                dt, de, db = dev_in[s]
                wf = f2.wf_device(idx, 'p', (dt, de, db))
                assert wf is not None, "the filtering library must expose its weights as `cl` (library_sepTP does)"
                G, C = qe.p(dt, de, db, *wf, merge_analysis=self.merge_analysis)
            comp_done[s].record(main)
            if pin_out[s] is None:
                pin_out[s] = [torch.empty(G.numel(), dtype=torch.complex128, pin_memory=True) for _ in range(2)]
            cs_out.wait_event(comp_done[s])
            with torch.cuda.stream(cs_out):
                pin_out[s][0].copy_(G, non_blocking=True)
                pin_out[s][1].copy_(C, non_blocking=True)
                d2h_done[s].record(cs_out)
            live[s].append((G, C))                       # device results stay allocated until their copy is done
            if prev is not None:
                yield finish(*prev)
            prev = (s, idx)
        if prev is not None:
            yield finish(*prev)

    # ---- GPU evaluation
    def _engine(self, lmax_ivf):
        if self._qe is None or self._qe.lmax_ivf != lmax_ivf:
            self._qe = qe_device(self.nside, lmax_ivf, self.lmax_qlm['T'])
        return self._qe

    def _legs(self, idx, k, swapped):
        f1 = self.f2map2 if swapped else self.f2map1
        f2 = self.f2map1 if swapped else self.f2map2
        return f1, f2

    # ---- device-resident evaluation: used whenever both filtering libraries hand out CUDA tensors
    #      (`get_sim_teblm_dev` / `get_sim_mliklm_dev`, filt_simple.library_sepTP and its wrappers): nothing of a
    #      simulation visits the host between the simulated map and the estimate
    def _dev_ok(self):
        return all(hasattr(f.ivfs, 'get_sim_teblm_dev') and hasattr(f.ivfs, 'get_sim_mliklm_dev')
                   for f in (self.f2map1, self.f2map2)) and os.environ.get('PLK_QE_DEVICE', '1') != '0'

    def _dev_gclm(self, idx, k, swapped=False, prefetch=()):
        """(G, C) CUDA tensors of a fundamental lensing key; same legs and per-l weights as the host-fed paths below"""
        f1, f2 = self._legs(idx, k, swapped)
        fields = {'ptt': 't', 'p_p': 'eb', 'p': 'teb'}[k]
        bars = dict(zip(fields, f1.ivfs.get_sim_teblm_dev(idx, fields)))
        if len(prefetch) and hasattr(f1.ivfs, 'prefetch_dev'):
            f1.ivfs.prefetch_dev(list(prefetch))        # the filters of the next simulations run under this estimate
        wfs = dict(zip(fields, f2.ivfs.get_sim_mliklm_dev(idx, fields)))
        lmax = sht.alm_lmax(next(iter(bars.values())).numel())
        qe = self._engine(lmax)
        if k == 'p':            # C^TE cross terms of the Wiener legs (qest.py:582-588, :613-618)
            b2 = bars if f2.ivfs is f1.ivfs else dict(zip('te', f2.ivfs.get_sim_teblm_dev(idx, 'te')))
            one, clte = f2._one_dev(lmax), f2._clte_dev(lmax)
            twf = sht.alm_combine([(wfs['t'], one), (b2['e'], clte)])
            ewf = sht.alm_combine([(wfs['e'], one), (b2['t'], clte)])
            return qe.p(bars['t'], bars['e'], bars['b'], twf, ewf, wfs['b'], merge_analysis=self.merge_analysis)
        if k == 'ptt':
            return qe.ptt(bars['t'], wfs['t'])
        return qe.p_p(bars['e'], bars['b'], wfs['e'], wfs['b'])

    def get_sim_qlm_dev(self, k, idx, prefetch=()):
        """Gradient and curl estimates of 'ptt', 'p_p' or 'p' as CUDA tensors, symmetrised like `get_sim_qlm` when the
        two legs differ; not cached (mean-field and spectra accumulations that stay on the GPU).

        prefetch: simulation indices the caller will ask for next, in order (a loop over simulations knows them):
        filtering libraries that can work ahead (`filt_simple.library_sepTP.prefetch_dev`) filter them while this
        estimate is evaluated."""
        assert k in ['ptt', 'p_p', 'p'], k
        assert self._dev_ok(), "both filtering libraries must provide device-resident alms"
        G, C = self._dev_gclm(idx, k, prefetch=prefetch)
        if not self.f2map1.ivfs == self.f2map2.ivfs:
            G2, C2 = self._dev_gclm(idx, k, swapped=True)
            G, C = (G + G2) * 0.5, (C + C2) * 0.5
        return G, C

    def _get_sim_Tgclm(self, idx, k, swapped=False):
        """T-only estimator, gradient and curl (reference: qest.py:248-263)."""
        if self._dev_ok():
            G, C = self._dev_gclm(idx, k, swapped)
            return self._engine(self._qe.lmax_ivf).to_host(G, C)
        f1, f2 = self._legs(idx, k, swapped)
        tbar = f1.ivfs.get_sim_tlm(idx)
        twf = f2.wf_tlm(idx, k)
        qe = self._engine(hp.Alm.getlmax(tbar.size))
        G, C = qe.ptt(sht.dev_alm(tbar), sht.dev_alm(twf))
        return G.cpu().numpy(), C.cpu().numpy()

    def _get_sim_Pgclm(self, idx, k, swapped=False):
        """P-only estimator (reference: qest.py:265-285)."""
        if self._dev_ok():
            G, C = self._dev_gclm(idx, k, swapped)
            return self._engine(self._qe.lmax_ivf).to_host(G, C)
        f1, f2 = self._legs(idx, k, swapped)
        ebar, bbar = f1.ivfs.get_sim_elm(idx), f1.ivfs.get_sim_blm(idx)
        ewf, bwf = f2.wf_eblm(idx, k)
        qe = self._engine(hp.Alm.getlmax(ebar.size))
        G, C = qe.p_p(sht.dev_alm(ebar), sht.dev_alm(bbar), sht.dev_alm(ewf), sht.dev_alm(bwf))
        return G.cpu().numpy(), C.cpu().numpy()

    def _get_sim_MVgclm(self, idx, k, swapped=False):
        """MV estimator = P + T pieces with the C^TE cross terms in the Wiener legs (reference: qest.py:318-322)."""
        assert k == 'p'
        if self._dev_ok():
            G, C = self._dev_gclm(idx, k, swapped)
            return self._engine(self._qe.lmax_ivf).to_host(G, C)
        f1, f2 = self._legs(idx, k, swapped)
        tbar, ebar, bbar = f1.ivfs.get_sim_tlm(idx), f1.ivfs.get_sim_elm(idx), f1.ivfs.get_sim_blm(idx)
        qe = self._engine(hp.Alm.getlmax(tbar.size))
        dt, de, db = sht.dev_alm(tbar), sht.dev_alm(ebar), sht.dev_alm(bbar)
        wf = f2.wf_device(idx, 'p', (dt, de, db) if f2 is f1 or f2.ivfs is f1.ivfs else None)
        if wf is None:
            twf = sht.dev_alm(f2.wf_tlm(idx, 'p'))
            ewf, bwf = [sht.dev_alm(a) for a in f2.wf_eblm(idx, 'p')]
        else:
            twf, ewf, bwf = wf
        G, C = qe.p(dt, de, db, twf, ewf, bwf, merge_analysis=self.merge_analysis)
        return qe.to_host(G, C)

    def _symmetrised(self, fun, idx, k):
        G, C = fun(idx, k)
        if not self.f2map1.ivfs == self.f2map2.ivfs:
            _G, _C = fun(idx, k, swapped=True)
            G = 0.5 * (G + _G)
            C = 0.5 * (C + _C)
        return G, C

    def _save(self, kg, kc, idx, G, C):
        _write_alm(os.path.join(self.lib_dir, 'sim_%s_%04d.fits' % (kg, idx) if idx != -1 else 'dat_%s.fits' % kg), G)
        _write_alm(os.path.join(self.lib_dir, 'sim_%s_%04d.fits' % (kc, idx) if idx != -1 else 'dat_%s.fits' % kc), C)

    def _build_sim_Tgclm(self, idx):
        self._save('ptt', 'xtt', idx, *self._symmetrised(self._get_sim_Tgclm, idx, 'ptt'))

    def _build_sim_Pgclm(self, idx):
        self._save('p_p', 'x_p', idx, *self._symmetrised(self._get_sim_Pgclm, idx, 'p_p'))

    def _build_sim_MVgclm(self, idx):
        self._save('p', 'x', idx, *self._symmetrised(self._get_sim_MVgclm, idx, 'p'))

    # ---- scalar-source estimators: products of two spin-0 / spin-2 real-space legs, one spin-0 analysis
    def _plans(self, lmax_ivf):
        return sht.get_plan(self.nside, lmax_ivf), sht.get_plan(self.nside, self.lmax_qlm['T'])

    def _scalar_qlm(self, prod, fac, lmax_ivf):
        fl = _dfl(fac * np.ones(self.lmax_qlm['T'] + 1))
        return self._plans(lmax_ivf)[1].map2alm(prod, fl=fl).cpu().numpy()

    def _t_res_map(self, f, idx, wl=None):
        """alm2map(w_l T_bar) on the device"""
        tlm = f.ivfs.get_sim_tlm(idx)
        lmax = hp.Alm.getlmax(tlm.size)
        return self._plans(lmax)[0].alm2map(sht.dev_alm(tlm), fl=None if wl is None else sht.dev_fl(wl, lmax)), lmax

    def _p_res_maps(self, f, idx):
        """alm2map_spin(E_bar / 2, B_bar / 2, 2)"""
        elm, blm = f.ivfs.get_sim_elm(idx), f.ivfs.get_sim_blm(idx)
        lmax = hp.Alm.getlmax(elm.size)
        half = _dfl(0.5 * np.ones(lmax + 1))
        return self._plans(lmax)[0].alm2map_spin(sht.dev_alm(elm), sht.dev_alm(blm), 2, flg=half, flc=half), lmax

    def _t_wf_map(self, f, idx, joint):
        tlm = f.ivfs.get_sim_tmliklm(idx)
        if joint:
            tlm = tlm + hp.almxfl(f.ivfs.get_sim_elm(idx), f.clte)
        lmax = hp.Alm.getlmax(tlm.size)
        return self._plans(lmax)[0].alm2map(sht.dev_alm(tlm))

    def _p_wf_maps(self, f, idx, joint):
        elm, blm = f.ivfs.get_sim_emliklm(idx), f.ivfs.get_sim_bmliklm(idx)
        if joint:
            elm = elm + hp.almxfl(f.ivfs.get_sim_tlm(idx), f.clte)
        lmax = hp.Alm.getlmax(elm.size)
        return self._plans(lmax)[0].alm2map_spin(sht.dev_alm(elm), sht.dev_alm(blm), 2)

    def _get_sim_stt(self, idx, swapped=False):
        """Point-source estimator -1/2 (T_bar_1 T_bar_2)_LM (reference: qest.py:286-290)."""
        f1, f2 = self._legs(idx, 'stt', swapped)
        t1, lmax = self._t_res_map(f1, idx)
        t2, _ = self._t_res_map(f2, idx)
        return self._scalar_qlm(t1.mul_(t2), -0.5, lmax)

    def _get_sim_ntt(self, idx, swapped=False):
        """Noise-inhomogeneity estimator: the same on beam-deconvolved maps (reference: qest.py:292-297)."""
        f1, f2 = self._legs(idx, 'ntt', swapped)
        t1, lmax = self._t_res_map(f1, idx, f1.ivfs.get_tal('t')[:])
        t2, _ = self._t_res_map(f2, idx, f2.ivfs.get_tal('t')[:])
        return self._scalar_qlm(t1.mul_(t2), -0.5, lmax)

    def _get_sim_ftt(self, idx, joint=False, swapped=False):
        """Modulation estimator, temperature: -(T_bar_1 T^WF_2)_LM (reference: qest.py:299-303)."""
        f1, f2 = self._legs(idx, 'ftt', swapped)
        t1, lmax = self._t_res_map(f1, idx)
        return self._scalar_qlm(t1.mul_(self._t_wf_map(f2, idx, joint)), -1.0, lmax)

    def _get_sim_f_p(self, idx, joint=False, swapped=False):
        """Modulation estimator, polarization: -2 (Q_1 Q_2 + U_1 U_2)_LM (reference: qest.py:305-309)."""
        f1, f2 = self._legs(idx, 'f_p', swapped)
        (Q1, U1), lmax = self._p_res_maps(f1, idx)
        Q2, U2 = self._p_wf_maps(f2, idx, joint)
        return self._scalar_qlm(Q1.mul_(Q2).add_(U1.mul_(U2)), -2.0, lmax)

    def _get_sim_a_p(self, idx, joint=False, swapped=False):
        """Polarization-rotation estimator: -4 (Q_1 U_2 - U_1 Q_2)_LM (reference: qest.py:311-315)."""
        f1, f2 = self._legs(idx, 'a_p', swapped)
        (Q1, U1), lmax = self._p_res_maps(f1, idx)
        Q2, U2 = self._p_wf_maps(f2, idx, joint)
        return self._scalar_qlm(Q1.mul_(U2).sub_(U1.mul_(Q2)), -4.0, lmax)

    def _two_legs(self):
        return not self.f2map1.ivfs == self.f2map2.ivfs

    def _save1(self, k, idx, qlm):
        _write_alm(os.path.join(self.lib_dir, 'sim_%s_%04d.fits' % (k, idx) if idx != -1 else 'dat_%s.fits' % k), qlm)

    def _build_sim_stt(self, idx):
        self._save1('stt', idx, self._get_sim_stt(idx))          # symmetric in the two legs: no swap needed

    def _build_sim_ntt(self, idx):
        self._save1('ntt', idx, self._get_sim_ntt(idx))

    def _build_sim_ftt(self, idx):
        f = self._get_sim_ftt(idx)
        if self._two_legs():
            f = 0.5 * (f + self._get_sim_ftt(idx, swapped=True))
        self._save1('ftt', idx, f)

    def _build_sim_f_p(self, idx):
        f = self._get_sim_f_p(idx)
        if self._two_legs():
            f = 0.5 * (f + self._get_sim_f_p(idx, swapped=True))
        self._save1('f_p', idx, f)

    def _build_sim_a_p(self, idx):
        a = self._get_sim_a_p(idx)
        if self._two_legs():
            # the reference averages with the swapped-leg *f_p* estimate here (qest.py:433); kept for parity
            a = 0.5 * (a + self._get_sim_f_p(idx, swapped=True))
        self._save1('a_p', idx, a)

    def _build_sim_f(self, idx):
        """MV modulation estimator: P and T pieces with the C^TE cross terms in the Wiener legs (reference: qest.py:359-367)."""
        fp = self._get_sim_f_p(idx, joint=True)
        ft = self._get_sim_ftt(idx, joint=True)
        if self._two_legs():
            fp = 0.5 * (fp + self._get_sim_f_p(idx, joint=True, swapped=True))
            ft = 0.5 * (ft + self._get_sim_ftt(idx, joint=True, swapped=True))
        self._save1('f', idx, fp + ft)

    # ---- single-pair lensing estimators 'pte', 'peb', ...: the MV estimator with one field kept on each leg
    def _get_sim_xfilt_gclm(self, idx, x1, x2, swapped=False):
        """MV lensing estimator keeping field x1 ('t', 'e' or 'b') on the inverse-variance leg and x2 on the
        Wiener-filtered gradient leg (reference: qest.py:369-399 with the xfilt branches of qest.py:506-638).
        Separately filtered libraries only, as in the reference."""
        f1, f2 = self._legs(idx, 'p', swapped)
        if swapped:
            x1, x2 = x2, x1
        assert isinstance(f2, lib_filt2map_sepTP), 'not implemented'
        bars = {'t': f1.ivfs.get_sim_tlm, 'e': f1.ivfs.get_sim_elm, 'b': f1.ivfs.get_sim_blm}
        bar = bars[x1](idx)
        lmax = hp.Alm.getlmax(bar.size)
        qe = self._engine(lmax)
        zero = torch.zeros(bar.size, dtype=torch.complex128, device='cuda')
        dbar = sht.dev_alm(bar)
        # Wiener legs sourced by x2 alone: T^WF = C^TT T_bar [x2 = t] + C^TE E_bar [x2 = e],
        #                                   E^WF = C^EE E_bar [x2 = e] + C^TE T_bar [x2 = t],  B^WF = C^BB B_bar [x2 = b]
        twf = ewf = bwf = None
        if x2 == 't':
            twf = sht.dev_alm(f2.ivfs.get_sim_tmliklm(idx))
            ewf = sht.dev_alm(hp.almxfl(f2.ivfs.get_sim_tlm(idx), f2.clte))
        elif x2 == 'e':
            twf = sht.dev_alm(hp.almxfl(f2.ivfs.get_sim_elm(idx), f2.clte))
            ewf = sht.dev_alm(f2.ivfs.get_sim_emliklm(idx))
        else:
            bwf = sht.dev_alm(f2.ivfs.get_sim_bmliklm(idx))
        re = im = None
        if x1 == 't' and twf is not None:
            G, C = qe.t_products(dbar, twf)
            re, im = G, C
        elif x1 in 'eb' and (ewf is not None or bwf is not None):
            re, im = qe.p_products(dbar if x1 == 'e' else zero, dbar if x1 == 'b' else zero,
                                   zero if ewf is None else ewf, zero if bwf is None else bwf)
        if re is None:
            n = hp.Alm.getsize(self.lmax_qlm['T'])
            return np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)
        return qe.to_host(*qe.analyse(re, im))

    def _build_sim_xfiltMVgclm(self, idx, k):
        assert k[0] in 'px' and k[1:] in _XY_PAIRS + ['tt'], k
        G, C = self._get_sim_xfilt_gclm(idx, k[-2], k[-1])
        if self._two_legs():
            _G, _C = self._get_sim_xfilt_gclm(idx, k[-2], k[-1], swapped=True)
            G, C = 0.5 * (G + _G), 0.5 * (C + _C)
        self._save('p' + k[1:], 'x' + k[1:], idx, G, C)


class lib_filt2map(object):
    """Filtered alms -> real-space legs, jointly filtered T and P (reference: qest.py:441-530).

    The `get_*map` methods return numpy maps as in the reference; `wf_tlm` / `wf_eblm` return the Wiener-filtered
    alms that enter the gradient legs (what the GPU path consumes)."""

    def __init__(self, ivfs, nside):
        self.ivfs = ivfs
        self.nside = nside

    def hashdict(self):
        return {'ivfs': self.ivfs.hashdict(), 'nside': self.nside}

    def wf_tlm(self, idx, k=None):
        return self.ivfs.get_sim_tmliklm(idx)

    def wf_eblm(self, idx, k=None):
        return self.ivfs.get_sim_emliklm(idx), self.ivfs.get_sim_bmliklm(idx)

    def wf_device(self, idx, k, dev_bars=None):
        """Wiener legs built on the GPU from the inverse-variance filtered alms, when the filtering library exposes
        its weights as `cl` (every `library_sepTP` does); None otherwise (host path through get_sim_*mliklm)."""
        return None

    def get_gtmap(self, idx, k=None, xfilt=None):
        r"""\sum_{lm} MAP_talm sqrt(l (l + 1)) _1 Ylm(n): spin-1 transform with zero curl (reference: qest.py:453-464)."""
        assert xfilt is None, 'not implemented'
        tlm = self.wf_tlm(idx, k)
        lmax = hp.Alm.getlmax(tlm.size)
        Glm = hp.almxfl(tlm, _grad_fl(lmax, 1, 't'))
        return hp.alm2map_spin([Glm, np.zeros_like(Glm)], self.nside, 1, lmax)

    def get_tmap(self, idx):
        return hp.alm2map(self.ivfs.get_sim_tmliklm(idx), self.nside)

    def get_pmap(self, idx):
        Glm, Clm = self.ivfs.get_sim_emliklm(idx), self.ivfs.get_sim_bmliklm(idx)
        return hp.alm2map_spin([Glm, Clm], self.nside, 2, hp.Alm.getlmax(Glm.size))

    def get_wirestmap(self, idx, wl):
        """weighted residual map alm2map(w_l T_bar) (reference: qest.py:516-519)"""
        reslm = self.ivfs.get_sim_tlm(idx)
        return hp.alm2map(hp.almxfl(reslm, wl), self.nside, lmax=hp.Alm.getlmax(reslm.size))

    def get_gpmap(self, idx, spin, k=None, xfilt=None):
        r"""\sum_{lm} (Elm +- iBlm) sqrt((l+2)(l-1)) _1 Ylm(n) or sqrt((l-2)(l+3)) _3 Ylm(n) (reference: qest.py:481-504)."""
        assert spin in [1, 3]
        assert xfilt is None, 'not implemented'
        Glm, Clm = self.wf_eblm(idx, k)
        lmax = hp.Alm.getlmax(Glm.size)
        fl = _grad_fl(lmax, spin, 'p')
        return hp.alm2map_spin([hp.almxfl(Glm, fl), hp.almxfl(Clm, fl)], self.nside, spin, lmax)

    def get_irestmap(self, idx, xfilt=None):
        assert xfilt is None, 'not implemented'
        reslm = self.ivfs.get_sim_tlm(idx)
        return hp.alm2map(reslm, self.nside, lmax=hp.Alm.getlmax(reslm.size))

    def get_irespmap(self, idx, xfilt=None):
        assert xfilt is None, 'not implemented'
        reselm, resblm = self.ivfs.get_sim_elm(idx), self.ivfs.get_sim_blm(idx)
        assert hp.Alm.getlmax(reselm.size) == hp.Alm.getlmax(resblm.size)
        return hp.alm2map_spin([reselm * 0.5, resblm * 0.5], self.nside, 2, hp.Alm.getlmax(reselm.size))


class lib_filt2map_sepTP(lib_filt2map):
    """Same for separately filtered T and P: the C^TE cross terms are added to the Wiener legs of the MV
    estimator (reference: qest.py:533-638)."""

    def __init__(self, ivfs, nside, clte):
        super(lib_filt2map_sepTP, self).__init__(ivfs, nside)
        self.clte = clte

    def hashdict(self):
        return {'ivfs': self.ivfs.hashdict(), 'nside': self.nside, 'clte': ut.clhash(self.clte)}

    def wf_tlm(self, idx, k=None):
        assert k in ['ptt', 'p'], k
        tlm = self.ivfs.get_sim_tmliklm(idx)
        if k == 'p':
            tlm = tlm + hp.almxfl(self.ivfs.get_sim_elm(idx), self.clte)      # qest.py:582-588
        return tlm

    def wf_eblm(self, idx, k=None):
        assert k in ['p_p', 'p'], k
        elm, blm = self.ivfs.get_sim_emliklm(idx), self.ivfs.get_sim_bmliklm(idx)
        if k == 'p':
            elm = elm + hp.almxfl(self.ivfs.get_sim_tlm(idx), self.clte)      # qest.py:613-618
        return elm, blm

    def _clte_dev(self, lmax):
        if not hasattr(self, '_clte_d'):
            self._clte_d = {}
        if lmax not in self._clte_d:
            self._clte_d[lmax] = sht.dev_fl(self.clte, lmax)
        return self._clte_d[lmax]

    def _one_dev(self, lmax):
        if not hasattr(self, '_one_d'):
            self._one_d = {}
        if lmax not in self._one_d:
            self._one_d[lmax] = torch.ones(lmax + 1, dtype=torch.float64, device='cuda')
        return self._one_d[lmax]

    def _cl_dev(self):
        """filtering weights C_l^{TT, EE, BB} of the ivfs and C_l^{TE}, on the device"""
        if not hasattr(self, '_cl_d'):
            cl = self.ivfs.cl
            self._cl_d = {x: _dfl(cl[x]) for x in ('tt', 'ee', 'bb')}
            self._cl_d['te'] = _dfl(self.clte)
        return self._cl_d

    def wf_device(self, idx, k, dev_bars=None):
        cl = getattr(self.ivfs, 'cl', None)
        if not isinstance(cl, dict) or not all(x in cl for x in ('tt', 'ee', 'bb')) or type(self.ivfs).__name__ in ('library_ftl', 'library_shuffle'):
            return None
        if dev_bars is None:
            dev_bars = [sht.dev_alm(a) for a in (self.ivfs.get_sim_tlm(idx), self.ivfs.get_sim_elm(idx), self.ivfs.get_sim_blm(idx))]
        dt, de, db = dev_bars
        c = self._cl_dev()
        lmax = sht.alm_lmax(dt.numel())
        if k == 'p':
            twf = _combine(lmax, [(dt, c['tt']), (de, c['te'])])       # qest.py:582-588
            ewf = _combine(lmax, [(de, c['ee']), (dt, c['te'])])       # qest.py:613-618
        else:
            twf = sht.almxfl(dt, c['tt'])
            ewf = sht.almxfl(de, c['ee'])
        bwf = sht.almxfl(db, c['bb'])
        return twf, ewf, bwf

    def get_tmap(self, idx, joint=False):
        tlm = self.ivfs.get_sim_tmliklm(idx)
        if joint:
            tlm = tlm + hp.almxfl(self.ivfs.get_sim_elm(idx), self.clte)
        return hp.alm2map(tlm, self.nside)

    def get_pmap(self, idx, joint=False):
        Glm, Clm = self.ivfs.get_sim_emliklm(idx), self.ivfs.get_sim_bmliklm(idx)
        if joint:
            Glm = Glm + hp.almxfl(self.ivfs.get_sim_tlm(idx), self.clte)
        return hp.alm2map_spin([Glm, Clm], self.nside, 2, hp.Alm.getlmax(Glm.size))
