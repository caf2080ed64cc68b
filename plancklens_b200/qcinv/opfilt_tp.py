r"""Joint temperature + polarization Wiener / inverse-variance filter operators on the GPU
(reference: plancklens/qcinv/opfilt_tp.py).

 :math:`S^{-1} (S^{-1} + Y^t N^{-1} Y)^{-1} Y^t N^{-1}` on (T, E, B) with the full per-l 3x3 signal covariance.

Plug-in surface of the reference module (`calc_prep`, `apply_fini`, `dot_op`, `fwd_op`, `pre_op_diag`,
`pre_op_dense`, `alm_filter_sinv`, `alm_filter_ninv`).  Vectors are `util_alm.teblm` of GPU-resident `dalm`s.
One `fwd_op`: a spin-0 and a spin-2 synthesis with the T / E / B transfer functions fused (what
`hp.alm2map(pol=True)` does in the reference, opfilt_tp.py:279), the N^{-1} per-pixel kernels (T with the
monopole / dipole / template-map projection of opfilt_tt, polarization scalar or QQ/QU/UU), the two analyses with
b_l npix/4pi fused, and one three-term per-l combination kernel per component for S^{-1} x.
"""
import ctypes

import numpy as np
import torch

from .. import hp, sht
from ..utils import clhash
from . import dense, opfilt_tt, util
from .util_alm import dalm, teblm


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def _as_teblm_dev(alm):
    """(teblm of dalm, converted?)"""
    if isinstance(alm.tlm, dalm):
        return alm, False
    return teblm([dalm.from_numpy(c) for c in (alm.tlm, alm.elm, alm.blm)]), True


def _combine(lmax, terms):
    """sum_j fl_j[l] a_j[l,m] for up to four (device alm tensor, device fl tensor) terms."""
    n = len(terms)
    out = torch.empty_like(terms[0][0])
    ins = (ctypes.c_void_p * n)(*[t[0].data_ptr() for t in terms])
    fls = (ctypes.c_void_p * n)(*[t[1].data_ptr() for t in terms])
    nfl = (ctypes.c_int * n)(*[int(t[1].numel()) for t in terms])
    sht.check(sht._lib.load().plk_alm_combine_dev(lmax, n, ins, fls, nfl, sht._ptr(out), sht._stream()))
    return out


class dot_op:
    """sum over T, E and B of sum_l (2l+1) C_l^{ab}, all multipoles  (reference: opfilt_tp.py:46-58)."""

    def __init__(self):
        pass

    def dev(self, alm1, alm2):
        assert alm1.lmaxt == alm2.lmaxt and alm1.lmaxe == alm2.lmaxe and alm1.lmaxb == alm2.lmaxb
        assert alm1.lmaxt == alm1.lmaxe == alm1.lmaxb
        return sht.alm_dotn([alm1.tlm.t, alm1.elm.t, alm1.blm.t], [alm2.tlm.t, alm2.elm.t, alm2.blm.t], lmin=0)

    def fused(self, alm1, alm2, num=None, den=None, scale=1.0):
        """[s, r, -r] on the device in one kernel (see opfilt_tt.dot_op.fused)"""
        assert alm1.lmaxt == alm1.lmaxe == alm1.lmaxb == alm2.lmaxt == alm2.lmaxe == alm2.lmaxb
        return sht.alm_dot_fused([alm1.tlm.t, alm1.elm.t, alm1.blm.t], [alm2.tlm.t, alm2.elm.t, alm2.blm.t], lmin=0,
                                 num=num, den=den, scale=scale)

    def __call__(self, alm1, alm2):
        return float(self.dev(alm1, alm2).item())


class _lmat3:
    """Per-l symmetric 3x3 matrix applied to a (T, E, B) triple: one combine kernel per component."""

    def __init__(self, mat):
        self.mat = mat
        self.te_only = not (np.any(mat[:, 0, 2]) or np.any(mat[:, 1, 2]))
        self._d = [[_dev(mat[:, i, j]) for j in range(3)] for i in range(3)]

    def apply(self, alm):
        lmax = alm.lmax
        t, e, b = alm.tlm.t, alm.elm.t, alm.blm.t
        d = self._d
        if self.te_only:
            rt = _combine(lmax, [(t, d[0][0]), (e, d[0][1])])
            re = _combine(lmax, [(t, d[1][0]), (e, d[1][1])])
            rb = sht.almxfl(b, d[2][2])
        else:
            rt = _combine(lmax, [(t, d[0][0]), (e, d[0][1]), (b, d[0][2])])
            re = _combine(lmax, [(t, d[1][0]), (e, d[1][1]), (b, d[1][2])])
            rb = _combine(lmax, [(t, d[2][0]), (e, d[2][1]), (b, d[2][2])])
        z = alm.is_zero()
        return teblm([dalm(rt, lmax, z), dalm(re, lmax, z), dalm(rb, lmax, z)])


class alm_filter_sinv:
    """Per-l pseudo-inverse of the TEB signal covariance (reference: opfilt_tp.py:126-163)."""

    def __init__(self, s_cls, lmax):
        slmat = np.zeros((lmax + 1, 3, 3))
        z = np.zeros(lmax + 1)
        slmat[:, 0, 0] = s_cls.get('tt', z)[:lmax + 1]
        slmat[:, 0, 1] = slmat[:, 1, 0] = s_cls.get('te', z)[:lmax + 1]
        slmat[:, 0, 2] = slmat[:, 2, 0] = s_cls.get('tb', z)[:lmax + 1]
        slmat[:, 1, 1] = s_cls.get('ee', z)[:lmax + 1]
        slmat[:, 1, 2] = slmat[:, 2, 1] = s_cls.get('eb', z)[:lmax + 1]
        slmat[:, 2, 2] = s_cls.get('bb', z)[:lmax + 1]
        self.lmax = lmax
        self.slinv = np.linalg.pinv(slmat)
        self.te_only = not (np.any(slmat[:, 0, 2]) or np.any(slmat[:, 1, 2]))
        self._op = None

    def calc(self, alm):
        if self._op is None:
            self._op = _lmat3(self.slinv)
            self._op.te_only = self.te_only
        a, host = _as_teblm_dev(alm)
        r = self._op.apply(a)
        return teblm(list(r.numpy())) if host else r

    def hashdict(self):
        return {'slinv': clhash(self.slinv.flatten())}


class fwd_op:
    """A x = S^{-1} x + B^t N^{-1} B x  (reference: opfilt_tp.py:61-82)."""

    def __init__(self, s_cls, n_inv_filt):
        lmax = len(n_inv_filt.b_transf) - 1
        self.s_inv_filt = alm_filter_sinv(s_cls, lmax)
        self.n_inv_filt = n_inv_filt

    def hashdict(self):
        return {'s_inv_filt': self.s_inv_filt.hashdict(), 'n_inv_filt': self.n_inv_filt.hashdict()}

    def __call__(self, alm):
        return self.calc(alm)

    def calc(self, alm):
        if alm.is_zero():     # A 0 = 0 exactly (see opfilt_pp.fwd_op)
            return alm * 1.0
        nlm = alm * 1.0
        self.n_inv_filt.apply_alm(nlm)
        slm = self.s_inv_filt.calc(alm)
        return nlm + slm


class pre_op_diag:
    """Per-l 3x3 preconditioner pinv(S^{-1} + diag(N_l^{-1}))  (reference: opfilt_tp.py:87-118)."""

    def __init__(self, s_cls, n_inv_filt):
        lmax = len(n_inv_filt.b_transf) - 1
        s_inv_filt = alm_filter_sinv(s_cls, lmax)
        assert (s_inv_filt.lmax + 1) >= len(n_inv_filt.b_transf)
        ninv_ftl, ninv_fel, ninv_fbl = n_inv_filt.get_ftebl()
        flmat = s_inv_filt.slinv[0:lmax + 1, :, :]
        flmat[:, 0, 0] += ninv_ftl
        flmat[:, 1, 1] += ninv_fel
        flmat[:, 2, 2] += ninv_fbl
        self.flmat = np.linalg.pinv(flmat)
        self.te_only = s_inv_filt.te_only
        self._op = _lmat3(self.flmat)
        self._op.te_only = self.te_only

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, alm):
        return self._op.apply(alm)


def pre_op_dense(lmax, fwd_op, cache_fname=None):
    return dense.pre_op_dense_tp(lmax, fwd_op, cache_fname=cache_fname)


class alm_filter_ninv(object):
    """Pixel-space inverse noise for (T, Q, U): n_inv = [TT, (QQ+UU)/2] or [TT, QQ, QU, UU], temperature templates
    (monopole, dipole, maps) projected out (reference: opfilt_tp.py:166-326)."""

    def __init__(self, n_inv, b_transf, b_transf_e=None, b_transf_b=None, marge_monopole=False, marge_dipole=False,
                 marge_maps_t=(), marge_maps_p=()):
        self.n_inv = []
        for tn in n_inv:
            if isinstance(tn, list):
                prod = np.asarray(util.read_map(tn[0]), dtype=float)
                for n in tn[1:]:
                    prod = prod * util.read_map(n)
                self.n_inv.append(np.asarray(prod, dtype=float))
            else:
                self.n_inv.append(np.asarray(util.read_map(tn), dtype=float))
        assert len(self.n_inv) in (2, 4), len(self.n_inv)
        npix = len(self.n_inv[0])
        for n in self.n_inv[1:]:
            assert len(n) == npix
        assert len(marge_maps_p) == 0
        self.b_transf_t = np.asarray(b_transf, dtype=float)
        self.b_transf_e = np.asarray(b_transf_e, dtype=float) if b_transf_e is not None else self.b_transf_t
        self.b_transf_b = np.asarray(b_transf_b, dtype=float) if b_transf_b is not None else self.b_transf_t
        assert len(self.b_transf_t) == len(self.b_transf_e) == len(self.b_transf_b)
        self.b_transf = (self.b_transf_t + self.b_transf_e + self.b_transf_t) / 3.      # as in the reference (:228)
        self.marge_monopole = marge_monopole
        self.marge_dipole = marge_dipole
        self._marge_maps_t = marge_maps_t
        self.npix = npix
        self.nside = hp.npix2nside(npix)
        # the temperature block (N^{-1} multiply + template projection) is exactly opfilt_tt's
        self._tfilt = opfilt_tt.alm_filter_ninv(self.n_inv[0], self.b_transf_t, marge_monopole=marge_monopole,
                                                marge_dipole=marge_dipole, marge_maps=marge_maps_t)
        self.templates_t = self._tfilt.templates
        self.templates_t_hash = [clhash(np.asarray(util.read_map(m), dtype=float)) for m in marge_maps_t]
        if len(self.templates_t) != 0:
            self.Pt_Nn1_P_inv = self._tfilt.Pt_Nn1_P_inv
        self.templates_p = []
        self._np_d = [_dev(n) for n in self.n_inv[1:]]
        self._fl_cache = {}

    def get_ftebl(self):
        s = lambda m: np.sum(m) / (4.0 * np.pi)
        if len(self.n_inv) == 2:
            npp = s(self.n_inv[1])
        else:
            npp = s(0.5 * (self.n_inv[1] + self.n_inv[3]))
        return s(self.n_inv[0]) * self.b_transf_t ** 2, npp * self.b_transf_e ** 2, npp * self.b_transf_b ** 2

    def hashdict(self):
        return {'n_inv': [clhash(n) for n in self.n_inv], 'b_transf': clhash(self.b_transf),
                'marge_monopole': self.marge_monopole, 'marge_dipole': self.marge_dipole,
                'templates_t_hash': self.templates_t_hash}

    def degrade(self, nside):
        if nside == self.nside:
            return self
        print("DEGRADING WITH NO MARGE MAPS")
        return alm_filter_ninv([sht.ud_grade_sum(sht.dev_map(n), nside).cpu().numpy() for n in self.n_inv], self.b_transf_t,
                               b_transf_e=self.b_transf_e, b_transf_b=self.b_transf_b,
                               marge_monopole=self.marge_monopole, marge_dipole=self.marge_dipole)

    def _fl(self, which, lmax):
        k = (which, lmax)
        if k not in self._fl_cache:
            b = {'t': self.b_transf_t, 'e': self.b_transf_e, 'b': self.b_transf_b}[which[0]]
            self._fl_cache[k] = sht.dev_fl(b * (self.npix / (4. * np.pi)) if which[1:] == 'out' else b, lmax)
        return self._fl_cache[k]

    def apply_alm(self, alm):
        """alm <- B^t N^{-1} B alm in place (reference: opfilt_tp.py:270-298)."""
        a, host = _as_teblm_dev(alm)
        lmax = a.lmax
        plan = sht.get_plan(self.nside, lmax)
        tmap = plan.alm2map(a.tlm.t, fl=self._fl('tin', lmax))
        qmap, umap = plan.alm2map_spin(a.elm.t, a.blm.t, 2, flg=self._fl('ein', lmax), flc=self._fl('bin', lmax))
        self.apply_map([tmap, qmap, umap])
        plan.map2alm(tmap, fl=self._fl('tout', lmax), out=a.tlm.t)
        plan.map2alm_spin(qmap, umap, 2, flg=self._fl('eout', lmax), flc=self._fl('bout', lmax), out=(a.elm.t, a.blm.t))
        a.tlm.zero = a.elm.zero = a.blm.zero = False
        if host:
            t, e, b = a.numpy()
            alm.tlm[:] = t
            alm.elm[:] = e
            alm.blm[:] = b

    def apply_map(self, amap):
        """(T, Q, U) <- N^{-1} (T, Q, U) with the temperature templates projected out, in place
        (reference: opfilt_tp.py:300-326)."""
        tmap, qmap, umap = amap
        host = not isinstance(tmap, torch.Tensor)
        t, q, u = [sht.dev_map(m) if host else m for m in (tmap, qmap, umap)]
        self._tfilt.apply_map(t)
        if len(self.n_inv) == 2:
            sht.map_mul2(q, u, self._np_d[0])
        else:
            sht.map_ninv3(q, u, self._np_d[0], self._np_d[1], self._np_d[2])
        if host:
            tmap[:] = t.cpu().numpy()
            qmap[:] = q.cpu().numpy()
            umap[:] = u.cpu().numpy()


def calc_prep(maps, s_cls, n_inv_filt):
    """b = B^t N^{-1} d for d = (T, Q, U)  (reference: opfilt_tp.py:14-31)."""
    ms = [m.clone() if isinstance(m, torch.Tensor) else sht.dev_map(util.read_map(m)) for m in maps]
    assert ms[0].numel() == ms[1].numel() == ms[2].numel()
    n_inv_filt.apply_map(ms)
    lmax = len(n_inv_filt.b_transf) - 1
    plan = sht.get_plan(n_inv_filt.nside, lmax)
    tlm = plan.map2alm(ms[0], fl=n_inv_filt._fl('tout', lmax))
    elm, blm = plan.map2alm_spin(ms[1], ms[2], 2, flg=n_inv_filt._fl('eout', lmax), flc=n_inv_filt._fl('bout', lmax))
    return teblm([dalm(tlm, lmax), dalm(elm, lmax), dalm(blm, lmax)])


def apply_fini(alm, s_cls, n_inv_filt):
    """Wiener solution -> inverse-variance filtered (T, E, B), in place (reference: opfilt_tp.py:34-40)."""
    lmax = len(n_inv_filt.b_transf) - 1
    ret = alm_filter_sinv(s_cls, lmax).calc(alm)
    for name in ('tlm', 'elm', 'blm'):
        dst, src = getattr(alm, name), getattr(ret, name)
        if isinstance(dst, dalm):
            dst.t.copy_(src.t)
        else:
            dst[:] = src


def apply_finiMLIK(alm, s_cls, n_inv_filt):
    """Keeps the Wiener-filtered (maximum-likelihood) solution (reference: opfilt_tp.py:43-44)."""
    pass
