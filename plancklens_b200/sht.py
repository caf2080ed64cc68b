"""Device-side transform plans: thin Python carrier around the C ABI (include/plk.h).

torch tensors are only carriers for device memory and streams; every FLOP runs in libplk_b200.so.
"""
import contextlib
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from ._lib import check, vp

_PLANS = {}
_TLS = threading.local()
MAX_LANES = 4


def lane():
    """lane of the calling thread (0 unless inside `use_lane`)"""
    return getattr(_TLS, 'lane', 0)


@contextlib.contextmanager
def use_lane(k):
    """Runs the enclosed calls of this thread in lane k: its own plans (`get_plan`: work buffers are per plan) and its own
    reduction scratch in the library, so that two host threads on two CUDA streams -- the T and the P filter of one
    simulation -- never share mutable device state.  Objects bind to a lane by wrapping their work in it (`filt_cinv`)."""
    k = int(k)
    assert 0 <= k < MAX_LANES
    old = lane()
    _TLS.lane = k
    check(_lib.load().plk_set_lane(k))
    try:
        yield
    finally:
        _TLS.lane = old
        check(_lib.load().plk_set_lane(old))


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.PlkError("plancklens_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_lmax(size):
    lmax = int(np.floor(np.sqrt(2 * size) - 1))
    assert (lmax + 1) * (lmax + 2) // 2 == size, size
    return lmax


def _ptr(t):
    return vp(t.data_ptr()) if t is not None else None


def _stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def dev_alm(a, device=None):
    """numpy / torch complex alm -> contiguous complex128 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        return a.to(device='cuda', dtype=torch.complex128).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).cuda()


def dev_map(m):
    if isinstance(m, torch.Tensor):
        return m.to(device='cuda', dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)).cuda()


def dev_fl(fl, lmax):
    """per-l factor padded / truncated to lmax+1 entries on the device (hp.almxfl zero-pads short fl)."""
    if fl is None:
        return None
    if isinstance(fl, torch.Tensor):
        fl = fl.detach().cpu().numpy()
    out = np.zeros(lmax + 1)
    n = min(lmax + 1, len(fl))
    out[:n] = np.asarray(fl, dtype=float)[:n]
    return torch.from_numpy(out).cuda()


class Plan:
    """One (nside, lmax) transform plan on the current CUDA device."""

    def __init__(self, nside, lmax):
        _require_cuda()
        self.lib = _lib.load()
        self.nside, self.lmax = int(nside), int(lmax)
        self.npix = 12 * self.nside ** 2
        self.nalm = alm_size(self.lmax)
        self.nring = 4 * self.nside - 1
        self.pitch = (self.lmax + 2) & ~1
        h = vp()
        check(self.lib.plk_plan_create(ctypes.byref(h), self.nside, self.lmax, self.lmax))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self.lib.plk_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def active_fraction(self, spin):
        """share of the (l, m, ring pair) volume the Legendre kernels walk (the rest is skipped near the poles)"""
        v = ctypes.c_double()
        check(self.lib.plk_plan_active_fraction(self._h, int(spin), ctypes.byref(v)))
        return float(v.value)

    def set_seed_threshold(self, exp2):
        """start threshold 2^exp2 of the Legendre recurrences (default 2^-60); seed tables are rebuilt on next use"""
        check(self.lib.plk_plan_set_seed_threshold(self._h, int(exp2)))

    def device_bytes(self):
        return int(self.lib.plk_plan_device_bytes(self._h))

    # ---- device tensors in / out
    def alm2map(self, alm, fl=None, out=None):
        assert alm.numel() == self.nalm, (alm.numel(), self.nalm)
        out = torch.empty(self.npix, dtype=torch.float64, device='cuda') if out is None else out
        check(self.lib.plk_alm2map_dev(self._h, 0, _ptr(alm), None, _ptr(fl), None, _ptr(out), None, _stream()))
        return out

    def alm2map_spin(self, glm, clm, spin, flg=None, flc=None, out=None):
        assert spin in (1, 2, 3), spin
        assert glm.numel() == self.nalm and (clm is None or clm.numel() == self.nalm)
        if out is None:
            out = (torch.empty(self.npix, dtype=torch.float64, device='cuda'),
                   torch.empty(self.npix, dtype=torch.float64, device='cuda'))
        check(self.lib.plk_alm2map_dev(self._h, spin, _ptr(glm), _ptr(clm), _ptr(flg), _ptr(flc),
                                       _ptr(out[0]), _ptr(out[1]), _stream()))
        return out

    def map2alm(self, m, fl=None, out=None):
        assert m.numel() == self.npix
        out = torch.empty(self.nalm, dtype=torch.complex128, device='cuda') if out is None else out
        check(self.lib.plk_map2alm_dev(self._h, 0, _ptr(m), None, _ptr(fl), None, _ptr(out), None, _stream()))
        return out

    def map2alm_add(self, m, fl, add, afl, out=None):
        """fl * map2alm(m) + afl[l] * add in one output pass (CG forward operators); add must not alias out"""
        assert m.numel() == self.npix and add.numel() == self.nalm and afl.numel() == self.lmax + 1
        out = torch.empty(self.nalm, dtype=torch.complex128, device='cuda') if out is None else out
        check(self.lib.plk_map2alm_add_dev(self._h, 0, _ptr(m), None, _ptr(fl), None, _ptr(add), _ptr(afl), None, None,
                                           _ptr(out), None, _stream()))
        return out

    def map2alm_spin_add(self, m1, m2, spin, flg, flc, addg, aflg, addc, aflc, out=None):
        assert spin in (1, 2, 3), spin
        assert addg.numel() == self.nalm and addc.numel() == self.nalm
        assert aflg.numel() == self.lmax + 1 and aflc.numel() == self.lmax + 1
        if out is None:
            out = (torch.empty(self.nalm, dtype=torch.complex128, device='cuda'),
                   torch.empty(self.nalm, dtype=torch.complex128, device='cuda'))
        check(self.lib.plk_map2alm_add_dev(self._h, spin, _ptr(m1), _ptr(m2), _ptr(flg), _ptr(flc), _ptr(addg), _ptr(aflg),
                                           _ptr(addc), _ptr(aflc), _ptr(out[0]), _ptr(out[1]), _stream()))
        return out

    @staticmethod
    def _pixprog(terms):
        """[(scale, a, b or None), ...] -> plk_pixprog: pixel value = sum scale * a[p] * b[p]"""
        assert 1 <= len(terms) <= _lib.PLK_MAX_PIX_TERMS, len(terms)
        q = _lib.PixProg()
        q.nterm = len(terms)
        for k, (s, a, b) in enumerate(terms):
            q.a[k] = a.data_ptr()
            q.b[k] = b.data_ptr() if b is not None else None
            q.scale[k] = float(s)
        return q

    def map2alm_pix(self, terms, fl=None, add=None, afl=None, out=None):
        """spin-0 analysis of the map sum_k scale_k a_k b_k, evaluated inside the ring kernel (never materialised)"""
        out = torch.empty(self.nalm, dtype=torch.complex128, device='cuda') if out is None else out
        q = self._pixprog(terms)
        check(self.lib.plk_map2alm_pix_dev(self._h, 0, ctypes.byref(q), None, _ptr(fl), None, _ptr(add), _ptr(afl), None, None,
                                           _ptr(out), None, _stream()))
        return out

    def map2alm_spin_pix(self, terms1, terms2, spin, flg=None, flc=None, addg=None, aflg=None, addc=None, aflc=None, out=None):
        """spin-s analysis of two maps given as pixel programs (QE leg products, N^-1 multiplies), optional additive term"""
        assert spin in (1, 2, 3), spin
        if out is None:
            out = (torch.empty(self.nalm, dtype=torch.complex128, device='cuda'),
                   torch.empty(self.nalm, dtype=torch.complex128, device='cuda'))
        q1, q2 = self._pixprog(terms1), self._pixprog(terms2)
        check(self.lib.plk_map2alm_pix_dev(self._h, spin, ctypes.byref(q1), ctypes.byref(q2), _ptr(flg), _ptr(flc), _ptr(addg),
                                           _ptr(aflg), _ptr(addc), _ptr(aflc), _ptr(out[0]), _ptr(out[1]), _stream()))
        return out

    def map2alm_spin(self, m1, m2, spin, flg=None, flc=None, out=None):
        assert spin in (1, 2, 3), spin
        assert m1.numel() == self.npix and m2.numel() == self.npix
        if out is None:
            out = (torch.empty(self.nalm, dtype=torch.complex128, device='cuda'),
                   torch.empty(self.nalm, dtype=torch.complex128, device='cuda'))
        check(self.lib.plk_map2alm_dev(self._h, spin, _ptr(m1), _ptr(m2), _ptr(flg), _ptr(flc),
                                       _ptr(out[0]), _ptr(out[1]), _stream()))
        return out

    # ---- stages on their own (profiling / tests)
    def new_phase(self):
        return torch.zeros((self.nring, self.pitch), dtype=torch.complex128, device='cuda')

    def legendre_synth(self, spin, a1, a2=None, fl1=None, fl2=None, X1=None, X2=None):
        X1 = self.new_phase() if X1 is None else X1
        X2 = (self.new_phase() if X2 is None else X2) if spin else None
        check(self.lib.plk_legendre_synth_dev(self._h, spin, _ptr(a1), _ptr(a2), _ptr(fl1), _ptr(fl2),
                                              _ptr(X1), _ptr(X2), _stream()))
        return X1, X2

    def legendre_anal(self, spin, X1, X2=None, fl1=None, fl2=None):
        a1 = torch.empty(self.nalm, dtype=torch.complex128, device='cuda')
        a2 = torch.empty(self.nalm, dtype=torch.complex128, device='cuda') if spin else None
        check(self.lib.plk_legendre_anal_dev(self._h, spin, _ptr(X1), _ptr(X2), _ptr(fl1), _ptr(fl2),
                                             _ptr(a1), _ptr(a2), _stream()))
        return a1, a2

    def ring_synth(self, X, out=None):
        out = torch.empty(self.npix, dtype=torch.float64, device='cuda') if out is None else out
        check(self.lib.plk_ring_synth_dev(self._h, _ptr(X), _ptr(out), _stream()))
        return out

    def ring_anal(self, m, X=None):
        X = self.new_phase() if X is None else X
        check(self.lib.plk_ring_anal_dev(self._h, _ptr(m), _ptr(X), _stream()))
        return X

    # ---- host (numpy) in / out through the library's own staging
    def alm2map_host(self, spin, a1, a2=None):
        a1 = np.ascontiguousarray(a1, dtype=np.complex128)
        assert a1.size == self.nalm, (a1.size, self.nalm)
        m1 = np.empty(self.npix)
        m2 = np.empty(self.npix) if spin else None
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, dtype=np.complex128)
        check(self.lib.plk_alm2map_host(self._h, spin, vp(a1.ctypes.data), vp(a2.ctypes.data) if a2 is not None else None,
                                        vp(m1.ctypes.data), vp(m2.ctypes.data) if spin else None))
        return (m1, m2) if spin else m1

    def map2alm_host(self, spin, m1, m2=None):
        m1 = np.ascontiguousarray(m1, dtype=np.float64)
        assert m1.size == self.npix
        a1 = np.empty(self.nalm, dtype=np.complex128)
        a2 = np.empty(self.nalm, dtype=np.complex128) if spin else None
        if spin:
            m2 = np.ascontiguousarray(m2, dtype=np.float64)
        check(self.lib.plk_map2alm_host(self._h, spin, vp(m1.ctypes.data), vp(m2.ctypes.data) if spin else None,
                                        vp(a1.ctypes.data), vp(a2.ctypes.data) if spin else None))
        return (a1, a2) if spin else a1

    # ---- template (monopole / dipole) passes
    def modes_dot(self, m, w=None, out=None):
        out = torch.empty(4, dtype=torch.float64, device='cuda') if out is None else out
        check(self.lib.plk_map_modes_dot_dev(self._h, _ptr(m), _ptr(w), _ptr(out), _stream()))
        return out

    def modes_sub(self, m, w, sums, pinv):
        check(self.lib.plk_map_modes_sub_dev(self._h, _ptr(m), _ptr(w), _ptr(sums), _ptr(pinv), _stream()))
        return m


def get_plan(nside, lmax):
    key = (int(nside), int(lmax), torch.cuda.current_device() if torch.cuda.is_available() else -1, lane())
    if key not in _PLANS:
        _PLANS[key] = Plan(nside, lmax)
    return _PLANS[key]


def clear_plans():
    _PLANS.clear()


# ---- alm BLAS-1 on device tensors
def almxfl(alm, fl, out=None):
    """hp.almxfl on a device alm; fl is a device float64 tensor (any length; l >= len(fl) -> 0)."""
    lib = _lib.load()
    lmax = alm_lmax(alm.numel())
    out = torch.empty_like(alm) if out is None else out
    check(lib.plk_almxfl_dev(lmax, _ptr(alm), _ptr(fl), int(fl.numel()), _ptr(out), _stream()))
    return out


def alm_axpy(y, x, a):
    """y += a x ; a is a python float or a 1-element device tensor."""
    lib = _lib.load()
    if isinstance(a, torch.Tensor):
        check(lib.plk_alm_axpy_dev(y.numel(), 0.0, _ptr(a), _ptr(x), _ptr(y), _stream()))
    else:
        check(lib.plk_alm_axpy_dev(y.numel(), float(a), None, _ptr(x), _ptr(y), _stream()))
    return y


def alm_dot(a, b, lmin=0, out=None):
    lib = _lib.load()
    lmax = alm_lmax(a.numel())
    out = torch.empty(1, dtype=torch.float64, device='cuda') if out is None else out
    check(lib.plk_alm_dot_dev(lmax, lmin, _ptr(a), _ptr(b), _ptr(out), _stream()))
    return out


def alm_dot2(a1, b1, a2, b2, lmin=0, out=None):
    """sum of the two weighted dot products (E and B of an eblm pair) as one device scalar"""
    lib = _lib.load()
    lmax = alm_lmax(a1.numel())
    out = torch.empty(1, dtype=torch.float64, device='cuda') if out is None else out
    check(lib.plk_alm_dot2_dev(lmax, lmin, _ptr(a1), _ptr(b1), _ptr(a2), _ptr(b2), _ptr(out), _stream()))
    return out


def alm_dotn(avec, bvec, lmin=0, out=None):
    """sum over components of the weighted dot products (teblm vectors) as one device scalar"""
    lib = _lib.load()
    n = len(avec)
    lmax = alm_lmax(avec[0].numel())
    out = torch.empty(1, dtype=torch.float64, device='cuda') if out is None else out
    pa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in avec])
    pb = (ctypes.c_void_p * n)(*[t.data_ptr() for t in bvec])
    check(lib.plk_alm_dotn_dev(lmax, lmin, n, pa, pb, _ptr(out), _stream()))
    return out


def alm_dot_fused(avec, bvec, lmin=0, num=None, den=None, scale=1.0, out=None):
    """One-kernel dot product of up to four alm pairs -> 3-element device tensor [s, r, -r] with r = scale * num / s
    (num given), scale * s / den (den given); num / den are 1-element device tensors (plk_alm_dot_fused_dev)"""
    n = len(avec)
    lmax = alm_lmax(avec[0].numel())
    out = torch.empty(3, dtype=torch.float64, device='cuda') if out is None else out
    pa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in avec])
    pb = (ctypes.c_void_p * n)(*[t.data_ptr() for t in bvec])
    check(_lib.load().plk_alm_dot_fused_dev(lmax, int(lmin), n, pa, pb, _ptr(num), _ptr(den), float(scale), _ptr(out), _stream()))
    return out


def alm_axpy2(y1, x1, y2, x2, a_dev):
    """y1 += a x1 ; y2 -= a x2, a a 1-element device tensor"""
    check(_lib.load().plk_alm_axpy2_dev(y1.numel(), _ptr(a_dev), _ptr(x1), _ptr(y1), _ptr(x2), _ptr(y2), _stream()))


def alm_combine(terms, out=None):
    """sum_j fl_j[l] a_j[l, m] on the device for up to four (alm tensor, device fl tensor) pairs (plk_alm_combine_dev)"""
    n = len(terms)
    lmax = alm_lmax(terms[0][0].numel())
    out = torch.empty_like(terms[0][0]) if out is None else out
    ins = (ctypes.c_void_p * n)(*[t[0].data_ptr() for t in terms])
    fls = (ctypes.c_void_p * n)(*[t[1].data_ptr() for t in terms])
    nfl = (ctypes.c_int * n)(*[int(t[1].numel()) for t in terms])
    check(_lib.load().plk_alm_combine_dev(lmax, n, ins, fls, nfl, _ptr(out), _stream()))
    return out


def alm2cl(a, b=None):
    """hp.alm2cl on device alms -> device float64[lmax + 1]"""
    b = a if b is None else b
    lmax = alm_lmax(a.numel())
    out = torch.empty(lmax + 1, dtype=torch.float64, device='cuda')
    check(_lib.load().plk_alm2cl_dev(lmax, _ptr(a), _ptr(b), _ptr(out), _stream()))
    return out


def scalar_ratio(num, den, scale=1.0, out=None):
    """scale * num / den on 1-element device tensors (no host synchronisation)"""
    out = torch.empty(1, dtype=torch.float64, device='cuda') if out is None else out
    check(_lib.load().plk_scalar_ratio_dev(_ptr(num), _ptr(den), float(scale), _ptr(out), _stream()))
    return out


def alm_copy(alm, lmax_out):
    lib = _lib.load()
    lmax_in = alm_lmax(alm.numel())
    out = torch.empty(alm_size(lmax_out), dtype=torch.complex128, device='cuda')
    check(lib.plk_alm_copy_dev(lmax_in, _ptr(alm), lmax_out, _ptr(out), _stream()))
    return out


def alm_splice(lo, hi, lsplit):
    lib = _lib.load()
    out = torch.empty_like(hi)
    check(lib.plk_alm_splice_dev(alm_lmax(lo.numel()), _ptr(lo), alm_lmax(hi.numel()), _ptr(hi), lsplit, _ptr(out), _stream()))
    return out


def alm_splice_xfl(lo, hi, fl, lsplit):
    """lo for l <= lsplit, fl[l] * hi above (one kernel; same numbers as almxfl followed by alm_splice)"""
    out = torch.empty_like(hi)
    check(_lib.load().plk_alm_splice_xfl_dev(alm_lmax(lo.numel()), _ptr(lo), alm_lmax(hi.numel()), _ptr(hi), _ptr(fl), int(fl.numel()),
                                            int(lsplit), _ptr(out), _stream()))
    return out


def map_axpy(y, x, a):
    """y += a x on real maps (even number of pixels)"""
    assert y.numel() % 2 == 0
    check(_lib.load().plk_alm_axpy_dev(y.numel() // 2, float(a), None, _ptr(x), _ptr(y), _stream()))
    return y


def profile_enable(on=True):
    check(_lib.load().plk_profile_enable(1 if on else 0))


def profile_read():
    """-> {kind: (count, total_ms)} for kinds synth0, synths, anal0, anals"""
    c = (ctypes.c_int * 5)()
    t = (ctypes.c_double * 5)()
    check(_lib.load().plk_profile_read(c, t))
    names = ['synth_spin0', 'synth_spins', 'anal_spin0', 'anal_spins', 'synth_grad']
    return {n: (int(c[i]), float(t[i])) for i, n in enumerate(names)}


def fp64_peak_tflops(reps=3):
    v = ctypes.c_double()
    check(_lib.load().plk_fp64_peak(ctypes.byref(v), reps))
    return float(v.value)


def ud_grade_sum(m, nside_out):
    """hp.ud_grade(m, nside_out, power=-2) of a RING map on the device: every coarse pixel = sum of its children"""
    nside_in = int(round((m.numel() // 12) ** 0.5))
    assert 12 * nside_in ** 2 == m.numel(), m.numel()
    out = torch.empty(12 * int(nside_out) ** 2, dtype=torch.float64, device='cuda')
    check(_lib.load().plk_udgrade_sum_dev(nside_in, _ptr(m), int(nside_out), _ptr(out), _stream()))
    return out


def map_mul(y, a):
    check(_lib.load().plk_map_mul_dev(y.numel(), _ptr(y), _ptr(a), _stream()))
    return y


def map_dot(a, b, out=None):
    """sum_p a_p b_p as a 1-element device tensor"""
    out = torch.empty(1, dtype=torch.float64, device='cuda') if out is None else out
    check(_lib.load().plk_map_dot_dev(a.numel(), _ptr(a), _ptr(b), _ptr(out), _stream()))
    return out


def map_axpy_dev(y, x, a_dev):
    """y += a x on real maps, a read from device memory (1-element tensor); even number of pixels"""
    assert y.numel() % 2 == 0
    check(_lib.load().plk_alm_axpy_dev(y.numel() // 2, 0.0, _ptr(a_dev), _ptr(x), _ptr(y), _stream()))
    return y


def map_mul2(g, c, t):
    check(_lib.load().plk_map_mul2_dev(g.numel(), _ptr(g), _ptr(c), _ptr(t), _stream()))


def map_qe_pp(q, u, g3, c3, g1, c1, re, im):
    check(_lib.load().plk_map_qe_pp_dev(q.numel(), _ptr(q), _ptr(u), _ptr(g3), _ptr(c3), _ptr(g1), _ptr(c1),
                                       _ptr(re), _ptr(im), _stream()))


def map_ninv3(q, u, nqq, nqu, nuu):
    check(_lib.load().plk_map_ninv3_dev(q.numel(), _ptr(q), _ptr(u), _ptr(nqq), _ptr(nqu), _ptr(nuu), _stream()))


def map_cmul_acc(ar, ai, br, bi, dr, di):
    """(dr + i di) += (ar + i ai)(br + i bi); ai / bi may be None"""
    check(_lib.load().plk_map_cmul_acc_dev(ar.numel(), _ptr(ar), _ptr(ai), _ptr(br), _ptr(bi), _ptr(dr), _ptr(di), _stream()))


# ---- counter-based Gaussian random numbers on the device (Philox4x32-10, plk_rng.cuh)
def randn(seed, stream_id, n, scale=1.0, add=None, out=None):
    """out[i] = (add[i] if add is given else 0) + scale * z_i with unit normals z_i that depend on (seed, stream_id, i)"""
    out = torch.empty(int(n), dtype=torch.float64, device='cuda') if out is None else out
    check(_lib.load().plk_randn_dev(int(seed), int(stream_id), int(n), float(scale), _ptr(add), _ptr(out), _stream()))
    return out


def randn_alm(seed, stream_id, lmax):
    """unit-variance alm phases of a real field (reference recipe sims/phas.py:162-168) as a complex128 CUDA tensor"""
    out = torch.empty(alm_size(lmax), dtype=torch.complex128, device='cuda')
    check(_lib.load().plk_randn_alm_dev(int(seed), int(stream_id), int(lmax), _ptr(out), _stream()))
    return out


def philox_words(seed, stream_id, ncalls):
    """raw generator output [ncalls, 4] (int32 carrier of the uint32 words); tests only"""
    out = torch.empty((int(ncalls), 4), dtype=torch.int32, device='cuda')
    check(_lib.load().plk_philox_words_dev(int(seed), int(stream_id), int(ncalls), _ptr(out), _stream()))
    return out
