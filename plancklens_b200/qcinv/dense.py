"""Dense low-l preconditioner (reference: plancklens/qcinv/dense.py).

The matrix of `fwd_op` in the real-harmonic basis is filled by applying the (GPU) operator to unit vectors,
eigen-decomposed once, and kept on the device; applying it is pack (alm2rlm) -> GEMV -> unpack kernels of
libplk_b200.  Cached on disk in the same [lmax, hashdict, minv] pickle the reference writes (dense.py:107-108).
"""
import os
import pickle as pk

import numpy as np
import torch

from .. import sht
from ..utils import enumerate_progress
from .util_alm import dalm, eblm, teblm


def alm2rlm(alm):
    """complex alm -> real-harmonic coefficients, host version (reference: dense.py:16-33)."""
    alm = np.asarray(alm)
    lmax = sht.alm_lmax(alm.size)
    rlm = np.zeros((lmax + 1) ** 2)
    ls = np.arange(lmax + 1)
    rlm[ls ** 2] = alm[ls].real
    for m in range(1, lmax + 1):
        blk = alm[m * (2 * lmax + 1 - m) // 2 + ls[m:]]
        rlm[ls[m:] ** 2 + 2 * m - 1] = blk.real * np.sqrt(2.)
        rlm[ls[m:] ** 2 + 2 * m] = blk.imag * np.sqrt(2.)
    return rlm


def rlm2alm(rlm):
    """inverse of alm2rlm, host version (reference: dense.py:36-53)."""
    lmax = int(np.sqrt(len(rlm)) - 1)
    assert (lmax + 1) ** 2 == len(rlm)
    alm = np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=complex)
    ls = np.arange(lmax + 1)
    alm[ls] = rlm[ls ** 2]
    for m in range(1, lmax + 1):
        alm[m * (2 * lmax + 1 - m) // 2 + ls[m:]] = (rlm[ls[m:] ** 2 + 2 * m - 1] + 1.j * rlm[ls[m:] ** 2 + 2 * m]) / np.sqrt(2.)
    return alm


def _d_alm2rlm(t, lmax, out):
    sht.check(sht._lib.load().plk_alm2rlm_dev(lmax, sht._ptr(t), sht._ptr(out), sht._stream()))


def _d_rlm2alm(r, lmax):
    out = torch.empty(sht.alm_size(lmax), dtype=torch.complex128, device='cuda')
    sht.check(sht._lib.load().plk_rlm2alm_dev(lmax, sht._ptr(r), sht._ptr(out), sht._stream()))
    return out


def _matvec(A, x):
    y = torch.empty_like(x)
    sht.check(sht._lib.load().plk_dense_matvec_dev(int(x.numel()), sht._ptr(A), sht._ptr(x), sht._ptr(y), sht._stream()))
    return y


class _pre_op_dense:
    ncomp = 1

    def __init__(self, lmax, fwd_op, cache_fname=None):
        self.lmax = lmax
        if cache_fname is not None and os.path.exists(cache_fname):
            with open(cache_fname, 'rb') as f:
                cache_lmax, cache_hashdict, cache_minv = pk.load(f)
            if lmax != cache_lmax or self.hashdict(lmax, fwd_op) != cache_hashdict:
                print("WARNING: PRE_OP_DENSE CACHE: hashcheck failed. recomputing.")
                os.remove(cache_fname)
                self.compute_minv(lmax, fwd_op, cache_fname=cache_fname)
            else:
                self.minv = cache_minv
        else:
            self.compute_minv(lmax, fwd_op, cache_fname=cache_fname)
        self._minv_d = torch.from_numpy(np.ascontiguousarray(self.minv)).cuda()

    # --- vector <-> packed real harmonics, device
    def _pack(self, v):
        n1 = (self.lmax + 1) ** 2
        r = torch.empty(self.ncomp * n1, dtype=torch.float64, device='cuda')
        for i, c in enumerate(self._comps(v)):
            _d_alm2rlm(c.t, self.lmax, r[i * n1:(i + 1) * n1])
        return r

    def _unpack(self, r):
        n1 = (self.lmax + 1) ** 2
        return self._wrap([dalm(_d_rlm2alm(r[i * n1:(i + 1) * n1], self.lmax), self.lmax) for i in range(self.ncomp)])

    def _ntmpl(self, fwd_op):
        raise NotImplementedError

    def compute_minv(self, lmax, fwd_op, cache_fname=None):
        self.lmax = lmax
        nrlm = self.ncomp * (lmax + 1) ** 2
        ntmpl = self._ntmpl(fwd_op)
        print("computing dense preconditioner:")
        print("     lmax  =", lmax)
        print("     ntmpl =", ntmpl)
        # Row i of the (symmetric) matrix = fwd_op applied to unit vector i.  The operator is a fixed, host-sync-free
        # launch sequence, so it is captured once as a CUDA graph and replayed nrlm times (4225 for the default T
        # chain): the fill is then bound by the small transforms, not by ~40 Python -> ctypes launches per row.
        tmat_d = torch.empty((nrlm, nrlm), dtype=torch.float64, device='cuda')
        unit = torch.zeros(nrlm, dtype=torch.float64, device='cuda')
        col = self._pack(fwd_op(self._unpack(unit + 1.0)))     # eager warm-up: plans, tables, per-l factors
        graph = None
        if os.environ.get('PLK_CG_GRAPH', '1') != '0':
            import gc
            graph = torch.cuda.CUDAGraph()
            gc.collect()
            gc_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                    col = self._pack(fwd_op(self._unpack(unit)))
            finally:
                if gc_on:
                    gc.enable()
        one = torch.ones(1, dtype=torch.float64, device='cuda')
        for j, i in enumerate_progress(np.arange(nrlm), label='filling matrix'):
            unit[i:i + 1].copy_(one)
            if graph is not None:
                graph.replay()
            else:
                col = self._pack(fwd_op(self._unpack(unit)))
            tmat_d[i].copy_(col)
            unit[i:i + 1].zero_()
        tmat = tmat_d.cpu().numpy().T.copy()   # rows were filled (contiguous copies): column i = A e_i as in the reference
        print("   inverting M...")
        eigv, eigw = np.linalg.eigh(tmat)
        assert np.all(eigv[ntmpl:] > 0.)
        eigv_inv = np.zeros_like(eigv)
        eigv_inv[ntmpl:] = 1.0 / eigv[ntmpl:]
        if ntmpl > 0:
            # the ntmpl lowest modes (marginalised templates, l < 2 in polarization) are left untouched
            print("     eigv[ntmpl-1] = ", eigv[ntmpl - 1])
            print("     eigv[ntmpl]   = ", eigv[ntmpl])
            eigv_inv[0:ntmpl] = 1.0
        self.minv = np.dot(np.dot(eigw, np.diag(eigv_inv)), np.transpose(eigw))
        if cache_fname is not None:
            # several ranks may build the same matrix at once (one process per GPU): write-then-rename keeps readers safe
            tmp = '%s.%d.tmp' % (cache_fname, os.getpid())
            with open(tmp, 'wb') as f:
                pk.dump([lmax, self.hashdict(lmax, fwd_op), self.minv], f)
            os.replace(tmp, cache_fname)

    @staticmethod
    def hashdict(lmax, fwd_op):
        return {'lmax': lmax, 'fwd_op': fwd_op.hashdict()}

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, talm):
        return self._unpack(_matvec(self._minv_d, self._pack(talm)))

    def calc_low(self, talm, lsplit):
        """the same on the l <= lsplit block of a longer single-component vector, packed without the truncating copy"""
        assert self.ncomp == 1 and lsplit == self.lmax and talm.lmax >= lsplit
        n1 = (self.lmax + 1) ** 2
        r = torch.empty(n1, dtype=torch.float64, device='cuda')
        sht.check(sht._lib.load().plk_alm2rlm_from_dev(self.lmax, talm.lmax, sht._ptr(talm.t), sht._ptr(r), sht._stream()))
        return self._unpack(_matvec(self._minv_d, r))


class pre_op_dense_tt(_pre_op_dense):
    """reference: dense.py:57-119"""
    ncomp = 1

    def _comps(self, v):
        return [v]

    def _wrap(self, comps):
        return comps[0]

    def _ntmpl(self, fwd_op):
        return int(sum(t.nmodes for t in fwd_op.n_inv_filt.templates))


pre_op_dense_kk = pre_op_dense_tt


class pre_op_dense_pp(_pre_op_dense):
    """reference: dense.py:123-202"""
    ncomp = 2

    def _comps(self, v):
        return [v.elm, v.blm]

    def _wrap(self, comps):
        return eblm(comps)

    def _ntmpl(self, fwd_op):
        n = 0
        tp = getattr(fwd_op.n_inv_filt, 'templates_p', None)
        if tp is not None:
            n += sum(t.nmodes for t in tp)
        return int(n + 8)   # (1 mono + 3 dip) * (e + b)

    @staticmethod
    def alm2rlm(alm):
        e, b = alm.numpy() if hasattr(alm, 'numpy') else (alm.elm, alm.blm)
        return np.concatenate([alm2rlm(e), alm2rlm(b)])

    @staticmethod
    def rlm2alm(rlm):
        n1 = len(rlm) // 2
        return eblm([rlm2alm(rlm[:n1]), rlm2alm(rlm[n1:])])


class pre_op_dense_tp(_pre_op_dense):
    """reference: dense.py:204-283"""
    ncomp = 3

    def _comps(self, v):
        return [v.tlm, v.elm, v.blm]

    def _wrap(self, comps):
        return teblm(comps)

    def _ntmpl(self, fwd_op):
        n = sum(t.nmodes for t in fwd_op.n_inv_filt.templates_t)     # includes mono and possibly dip
        n += sum(t.nmodes for t in fwd_op.n_inv_filt.templates_p)
        return int(n + 8)   # (1 mono + 3 dip) * (e + b)

    @staticmethod
    def alm2rlm(alm):
        t, e, b = alm.numpy() if hasattr(alm, 'numpy') else (alm.tlm, alm.elm, alm.blm)
        return np.concatenate([alm2rlm(t), alm2rlm(e), alm2rlm(b)])

    @staticmethod
    def rlm2alm(rlm):
        n1 = len(rlm) // 3
        return teblm([rlm2alm(rlm[:n1]), rlm2alm(rlm[n1:2 * n1]), rlm2alm(rlm[2 * n1:])])
