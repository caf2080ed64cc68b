r"""Polarization-only Wiener / inverse-variance filter operators on the GPU (reference: plancklens/qcinv/opfilt_pp.py).

 :math:`S^{-1} (S^{-1} + Y^t N^{-1} Y)^{-1} Y^t N^{-1}`

Plug-in surface of the reference module (`calc_prep`, `apply_fini`, `dot_op`, `fwd_op`, `pre_op_diag`,
`pre_op_dense`, `alm_filter_sinv`, `alm_filter_ninv`).  Vectors are `util_alm.eblm` of GPU-resident `dalm`s.
One `fwd_op`: spin-2 synthesis with the E/B transfer functions fused, the N^{-1} per-pixel kernel
(scalar, or the symmetric QQ/QU/UU form), spin-2 analysis with b_l npix/4pi fused, and one four-term per-l
combination kernel per component for S^{-1} x + N x.
"""
import ctypes
import os

import numpy as np
import torch

from .. import hp, sht
from ..utils import clhash
from . import dense, template_removal, util
from .util_alm import dalm, eblm


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def _as_eblm_dev(alm):
    """(eblm of dalm, converted?)"""
    if isinstance(alm.elm, dalm):
        return alm, False
    return eblm([dalm.from_numpy(alm.elm), dalm.from_numpy(alm.blm)]), True


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
    """sum over E and B of sum_{l>=2} (2l+1) C_l^{ab}  (reference: opfilt_pp.py:27-34)."""

    def __init__(self):
        pass

    def __call__(self, alm1, alm2):
        assert alm1.lmax == alm2.lmax
        e = sht.alm_dot(alm1.elm.t, alm2.elm.t, lmin=2)
        b = sht.alm_dot(alm1.blm.t, alm2.blm.t, lmin=2)
        return float(e.item()) + float(b.item())

    def dev(self, alm1, alm2):
        """same number as a 1-element device tensor: no host synchronisation (fixed-iteration multigrid stages)"""
        assert alm1.lmax == alm2.lmax
        return sht.alm_dot2(alm1.elm.t, alm2.elm.t, alm1.blm.t, alm2.blm.t, lmin=2)

    def fused(self, alm1, alm2, num=None, den=None, scale=1.0):
        """[s, r, -r] on the device in one kernel (see opfilt_tt.dot_op.fused)"""
        assert alm1.lmax == alm2.lmax
        return sht.alm_dot_fused([alm1.elm.t, alm1.blm.t], [alm2.elm.t, alm2.blm.t], lmin=2, num=num, den=den, scale=scale)


class _lmat2:
    """Per-l symmetric 2x2 matrix applied to an (E, B) pair: one four-term combine per component."""

    def __init__(self, mat):
        self.mat = mat
        self._d = [[_dev(mat[:, i, j]) for j in range(2)] for i in range(2)]

    def apply(self, alm):
        lmax = alm.lmax
        e, b = alm.elm.t, alm.blm.t
        re = _combine(lmax, [(e, self._d[0][0]), (b, self._d[0][1])])
        rb = _combine(lmax, [(e, self._d[1][0]), (b, self._d[1][1])])
        z = alm.is_zero()
        return eblm([dalm(re, lmax, z), dalm(rb, lmax, z)])


class alm_filter_sinv:
    """Per-l pseudo-inverse of the (EE, EB; EB, BB) signal matrix (reference: opfilt_pp.py:87-107)."""

    def __init__(self, s_cls, lmax):
        slmat = np.zeros((lmax + 1, 2, 2), dtype=float)
        slmat[:, 0, 0] = s_cls.get('ee', np.zeros(lmax + 1))[:lmax + 1]
        slmat[:, 0, 1] = s_cls.get('eb', np.zeros(lmax + 1))[:lmax + 1]
        slmat[:, 1, 0] = s_cls.get('eb', np.zeros(lmax + 1))[:lmax + 1]
        slmat[:, 1, 1] = s_cls.get('bb', np.zeros(lmax + 1))[:lmax + 1]
        slinv = np.linalg.pinv(slmat)       # stacked: bit-identical to the per-l loop of the reference, 20x faster
        self.lmax = lmax
        self.slinv = slinv
        self._op = None

    def calc(self, alm):
        if self._op is None:
            self._op = _lmat2(self.slinv)
        a, host = _as_eblm_dev(alm)
        r = self._op.apply(a)
        return eblm(list(r.numpy())) if host else r

    def hashdict(self):
        return {'slinv': clhash(self.slinv.flatten())}


class fwd_op:
    """A x = S^{-1} x + B^t N^{-1} B x  (reference: opfilt_pp.py:37-55; no zero short-circuit there either)."""

    def __init__(self, s_cls, n_inv_filt):
        lmax = len(n_inv_filt.b_transf) - 1
        self.s_inv_filt = alm_filter_sinv(s_cls, lmax)
        self.n_inv_filt = n_inv_filt
        self._sl_d = {}

    def hashdict(self):
        return {'s_inv_filt': self.s_inv_filt.hashdict(), 'n_inv_filt': self.n_inv_filt.hashdict()}

    def __call__(self, alm):
        return self.calc(alm)

    def calc(self, alm):
        if alm.is_zero():
            # A 0 = 0 exactly (every stage is linear): skip the two spin-2 transforms on the zero start vector of the
            # solver.  The reference has this shortcut in opfilt_tt (opfilt_tt.py:68) only; the result is identical.
            return alm * 1.0
        sl = self.s_inv_filt.slinv
        if isinstance(alm.elm, dalm) and not np.any(sl[:, 0, 1]):
            # diagonal S^-1 (no EB spectrum, the usual case): apply_alm (opfilt_pp.py:253-270) with the S^-1 x term
            # folded into the output pass of the spin-2 analysis -- two transforms and the N^-1 kernel, nothing else
            nf = self.n_inv_filt
            nf._load_ninv()
            lmax = alm.lmax
            plan = sht.get_plan(nf.nside, lmax)
            qmap, umap = plan.alm2map_spin(alm.elm.t, alm.blm.t, 2, flg=nf._fl('ein', lmax), flc=nf._fl('bin', lmax))
            if lmax not in self._sl_d:
                self._sl_d[lmax] = (sht.dev_fl(sl[:, 0, 0], lmax), sht.dev_fl(sl[:, 1, 1], lmax))
            se, sb = self._sl_d[lmax]
            if len(nf.n_inv) == 1 and not nf.wmarg and os.environ.get('PLK_CG_PIXFUSED', '1') != '0':
                # one N^-1 map, no templates (opfilt_pp.py:292): the multiply happens inside the analysis ring kernel
                ninv = nf._ninv_d[0]
                e, b = plan.map2alm_spin_pix([(1.0, qmap, ninv)], [(1.0, umap, ninv)], 2, flg=nf._fl('eout', lmax),
                                             flc=nf._fl('bout', lmax), addg=alm.elm.t, aflg=se, addc=alm.blm.t, aflc=sb)
            else:
                nf.apply_map([qmap, umap])
                e, b = plan.map2alm_spin_add(qmap, umap, 2, nf._fl('eout', lmax), nf._fl('bout', lmax), alm.elm.t, se, alm.blm.t, sb)
            return eblm([dalm(e, lmax), dalm(b, lmax)])
        nlm = alm * 1.0
        self.n_inv_filt.apply_alm(nlm)
        slm = self.s_inv_filt.calc(alm)
        return nlm + slm


class pre_op_diag:
    """Per-l 2x2 preconditioner pinv(S^{-1} + diag(b_e^2, b_b^2)/N_l)  (reference: opfilt_pp.py:57-80)."""

    def __init__(self, s_cls, n_inv_filt):
        lmax = len(n_inv_filt.b_transf) - 1
        s_inv_filt = alm_filter_sinv(s_cls, lmax)
        assert (s_inv_filt.lmax + 1) >= len(n_inv_filt.b_transf)
        ninv_fel, ninv_fbl = n_inv_filt.get_febl()
        flmat = s_inv_filt.slinv
        flmat[:, 0, 0] += ninv_fel[:lmax + 1]
        flmat[:, 1, 1] += ninv_fbl[:lmax + 1]
        self.flmat = np.linalg.pinv(flmat)
        self._op = _lmat2(self.flmat)

    def __call__(self, talm):
        return self.calc(talm)

    def calc(self, alm):
        return self._op.apply(alm)


def pre_op_dense(lmax, fwd_op, cache_fname=None):
    return dense.pre_op_dense_pp(lmax, fwd_op, cache_fname=cache_fname)


class alm_filter_ninv(object):
    """Pixel-space polarization inverse noise: one map (QQ = UU) or three (QQ, QU, UU)
    (reference: opfilt_pp.py:110-303), with independent Q and U template marginalisation (marge_qmaps /
    marge_umaps, single n_inv map only -- as in the reference, opfilt_pp.py:153)."""

    def __init__(self, n_inv, b_transf, nlev_febl=None, b_transf_b=None, marge_qmaps=(), marge_umaps=()):
        self.b_transf_e = np.asarray(b_transf, dtype=float)
        self.b_transf_b = np.asarray(b_transf_b, dtype=float) if b_transf_b is not None else self.b_transf_e
        self.b_transf = 0.5 * (self.b_transf_e + self.b_transf_b)
        self.nside = None
        self.n_inv = None
        self.nlev_febl = nlev_febl
        self._n_inv = n_inv   # maps, paths or lists of those
        self.marge_qmaps = marge_qmaps
        self.marge_umaps = marge_umaps
        self.wmarg = (max(len(self.marge_qmaps), len(self.marge_umaps)) > 0)
        self.tniti = None
        self.templates_p = []
        self._ninv_d = None
        self._fl_cache = {}

    def _load_ninv(self):
        if self.n_inv is None:
            self.n_inv = []
            for tn in self._n_inv:
                self.n_inv.append(np.asarray(util.read_map(tn), dtype=float))
            assert len(self.n_inv) in [1, 3], len(self.n_inv)
            self.nside = hp.npix2nside(len(self.n_inv[0]))
            self._ninv_d = [_dev(n) for n in self.n_inv]

    def _build_tniti(self):
        """(P^t N^{-1} P)^{-1} of the Q and of the U templates, block diagonal (reference: opfilt_pp.py:149-185)."""
        if not self.wmarg or self.tniti is not None:
            return
        self._load_ninv()
        assert len(self.n_inv) == 1, 'QQ QU UU not implemented'
        blocks = []
        for im, marge_m in enumerate((self.marge_qmaps, self.marge_umaps)):
            if len(marge_m) == 0:
                continue
            tfunc = template_removal.template_qmap if im == 0 else template_removal.template_umap
            templates = [tfunc(np.asarray(util.read_map(m), dtype=float)) for m in marge_m]
            ntm = [sht.map_mul(t.map.clone(), self._ninv_d[0]) for t in templates]
            n = len(templates)
            mat = np.zeros((n, n))
            for i in range(n):
                for j in range(i + 1):
                    mat[i, j] = mat[j, i] = float(sht.map_dot(templates[j].map, ntm[i]).item())
            eigv, eigw = np.linalg.eigh(mat)
            blocks.append(np.dot(np.dot(eigw, np.diag(1.0 / eigv)), np.transpose(eigw)))
            self.templates_p += templates
            self._ntm_p = getattr(self, '_ntm_p', []) + ntm
        nmodes = sum(b.shape[0] for b in blocks)
        self.tniti = np.zeros((nmodes, nmodes))
        i0 = 0
        for b in blocks:
            self.tniti[i0:i0 + b.shape[0], i0:i0 + b.shape[0]] = b
            i0 += b.shape[0]
        self._mtniti_d = _dev(-self.tniti.reshape(-1))
        self._csum = torch.zeros(nmodes, dtype=torch.float64, device='cuda')
        self._mcoef = torch.zeros(nmodes, dtype=torch.float64, device='cuda')

    def _calc_febl(self):
        self._load_ninv()
        if len(self.n_inv) == 1:
            nlev_febl = 10800. / np.sqrt(np.sum(self.n_inv[0]) / (4.0 * np.pi)) / np.pi
        else:
            nlev_febl = 10800. / np.sqrt(np.sum(0.5 * (self.n_inv[0] + self.n_inv[2])) / (4.0 * np.pi)) / np.pi
        print("ninv_febl: using %.2f uK-amin noise Cl" % nlev_febl)
        return nlev_febl

    def get_ninv(self):
        self._load_ninv()
        return self.n_inv

    def get_mask(self):
        ninv = self.get_ninv()
        mask = np.where(ninv[0] > 0, 1., 0)
        for ni in ninv[1:]:
            mask *= (ni > 0)
        return mask

    def get_febl(self):
        if self.nlev_febl is None:
            self.nlev_febl = self._calc_febl()
        n_inv_cl_e = self.b_transf_e ** 2 / (self.nlev_febl / 180. / 60. * np.pi) ** 2
        n_inv_cl_b = self.b_transf_b ** 2 / (self.nlev_febl / 180. / 60. * np.pi) ** 2
        return n_inv_cl_e, n_inv_cl_b

    def hashdict(self):
        t_hash = []
        if self.wmarg:
            t_hash = [util.mask_hash(m, dtype=np.float32) for m in self.marge_qmaps]
            t_hash += [util.mask_hash(m, dtype=np.float32) for m in self.marge_umaps]
        return {'n_inv': [util.mask_hash(n, dtype=np.float16) for n in self._n_inv],
                'b_transf': clhash(self.b_transf), 'templates_p': t_hash}

    def degrade(self, nside):
        self._load_ninv()
        if nside == self.nside:
            return self
        return alm_filter_ninv([sht.ud_grade_sum(sht.dev_map(n), nside).cpu().numpy() for n in self.n_inv], self.b_transf_e,
                               b_transf_b=self.b_transf_b)

    def _fl(self, which, lmax):
        k = (which, lmax)
        if k not in self._fl_cache:
            npix = len(self.n_inv[0])
            b = self.b_transf_e if which[0] == 'e' else self.b_transf_b
            self._fl_cache[k] = sht.dev_fl(b * (npix / (4. * np.pi)) if which[1:] == 'out' else b, lmax)
        return self._fl_cache[k]

    def apply_alm(self, alm):
        """alm <- B^t N^{-1} B alm in place (reference: opfilt_pp.py:253-270)."""
        self._load_ninv()
        a, host = _as_eblm_dev(alm)
        lmax = a.lmax
        plan = sht.get_plan(self.nside, lmax)
        qmap, umap = plan.alm2map_spin(a.elm.t, a.blm.t, 2, flg=self._fl('ein', lmax), flc=self._fl('bin', lmax))
        self.apply_map([qmap, umap])
        plan.map2alm_spin(qmap, umap, 2, flg=self._fl('eout', lmax), flc=self._fl('bout', lmax),
                          out=(a.elm.t, a.blm.t))
        a.elm.zero = a.blm.zero = False
        if host:
            e, b = a.numpy()
            alm.elm[:] = e
            alm.blm[:] = b

    def apply_map(self, amap):
        """(Q, U) <- N^{-1} (Q, U) in place (reference: opfilt_pp.py:272-303)."""
        self._load_ninv()
        qmap, umap = amap
        host = not isinstance(qmap, torch.Tensor)
        q = sht.dev_map(qmap) if host else qmap
        u = sht.dev_map(umap) if host else umap
        if len(self.n_inv) == 1:
            sht.map_mul2(q, u, self._ninv_d[0])
            if self.wmarg:
                # (q, u) -= N^{-1} P (P^t N^{-1} P)^{-1} P^t (q, u)   (reference: opfilt_pp.py:279-290), on the device
                self._build_tniti()
                for j, t in enumerate(self.templates_p):
                    sht.map_dot(t.map, q if t.comp == 0 else u, out=self._csum[j:j + 1])
                n = self._csum.numel()
                sht.check(sht._lib.load().plk_dense_matvec_dev(n, sht._ptr(self._mtniti_d), sht._ptr(self._csum),
                                                               sht._ptr(self._mcoef), sht._stream()))
                for j, t in enumerate(self.templates_p):
                    sht.map_axpy_dev(q if t.comp == 0 else u, self._ntm_p[j], self._mcoef[j:j + 1])
        else:
            sht.map_ninv3(q, u, self._ninv_d[0], self._ninv_d[1], self._ninv_d[2])
        if host:
            qmap[:] = q.cpu().numpy()
            umap[:] = u.cpu().numpy()


def calc_prep(maps, s_cls, n_inv_filt):
    """b = B^t N^{-1} d for d = (Q, U)  (reference: opfilt_pp.py:306-317)."""
    qmap = sht.dev_map(util.read_map(maps[0])) if not isinstance(maps[0], torch.Tensor) else maps[0].clone()
    umap = sht.dev_map(util.read_map(maps[1])) if not isinstance(maps[1], torch.Tensor) else maps[1].clone()
    assert qmap.numel() == umap.numel()
    lmax = len(n_inv_filt.b_transf) - 1
    n_inv_filt.apply_map([qmap, umap])
    plan = sht.get_plan(n_inv_filt.nside, lmax)
    elm, blm = plan.map2alm_spin(qmap, umap, 2, flg=n_inv_filt._fl('eout', lmax), flc=n_inv_filt._fl('bout', lmax))
    return eblm([dalm(elm, lmax), dalm(blm, lmax)])


_FINI_CACHE = {}


def apply_fini(alm, s_cls, n_inv_filt):
    """Wiener solution -> inverse-variance filtered (E, B), in place (reference: opfilt_pp.py:320-324)."""
    # one filter object (per-l pseudo-inverses + device tables) per (spectra, lmax): built once, not once per solve
    key = (id(s_cls), alm.lmax)
    if key not in _FINI_CACHE:
        _FINI_CACHE[key] = (s_cls, alm_filter_sinv(s_cls, alm.lmax))      # s_cls kept alive: its id cannot be reused
    sfilt = _FINI_CACHE[key][1]
    ret = sfilt.calc(alm)
    if isinstance(alm.elm, dalm):
        alm.elm.t.copy_(ret.elm.t)
        alm.blm.t.copy_(ret.blm.t)
    else:
        alm.elm[:] = ret.elm
        alm.blm[:] = ret.blm
