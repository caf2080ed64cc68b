"""alm containers for the CG solver (reference: plancklens/qcinv/util_alm.py).

`alm_copy` / `alm_splice` accept numpy arrays (host, as in the reference) or the device containers below,
which expose methods of the same names -- the dispatch-on-attribute idiom the reference uses (util_alm.py:12, :31).
"""
import numpy as np
import torch

from .. import sht


def _getlmax(n):
    return sht.alm_lmax(n)


def alm_splice(alm_lo, alm_hi, lsplit):
    """alm with lmax(alm_hi): alm_lo for l <= lsplit, alm_hi above (reference: util_alm.py:8)."""
    if hasattr(alm_lo, 'alm_splice'):
        return alm_lo.alm_splice(alm_hi, lsplit)
    lo_lmax, hi_lmax = _getlmax(len(alm_lo)), _getlmax(len(alm_hi))
    assert lo_lmax >= lsplit and hi_lmax >= lsplit
    out = np.copy(alm_hi)
    for m in range(lsplit + 1):
        oh = m * (2 * hi_lmax + 1 - m) // 2
        ol = m * (2 * lo_lmax + 1 - m) // 2
        out[oh + m:oh + lsplit + 1] = alm_lo[ol + m:ol + lsplit + 1]
    return out


def alm_copy(alm, lmax=None):
    """Copy, optionally truncated to lmax (reference: util_alm.py:27)."""
    if hasattr(alm, 'alm_copy'):
        return alm.alm_copy(lmax=lmax)
    lmox = _getlmax(len(alm))
    assert lmax is None or lmax <= lmox
    if lmax is None or lmax == lmox:
        return np.copy(alm)
    out = np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=complex)
    for m in range(lmax + 1):
        o = m * (2 * lmax + 1 - m) // 2
        i = m * (2 * lmox + 1 - m) // 2
        out[o + m:o + lmax + 1] = alm[i + m:i + lmax + 1]
    return out


class dalm:
    """One alm vector resident on the GPU (complex128, healpy layout) with the arithmetic cd_solve needs.

    All arithmetic runs in libplk_b200 kernels (plk_alm_lincomb_dev / plk_alm_axpy_dev); torch only owns memory.
    `zero` is a host-side flag that stays True while the vector is known to be identically zero, which is what
    `opfilt_tt.fwd_op.calc` tests before doing any work (reference: opfilt_tt.py:68).
    """
    __slots__ = ('t', 'lmax', 'zero')

    def __init__(self, t, lmax=None, zero=False):
        self.t = t
        self.lmax = sht.alm_lmax(t.numel()) if lmax is None else lmax
        self.zero = zero

    @staticmethod
    def zeros(lmax):
        return dalm(torch.zeros(sht.alm_size(lmax), dtype=torch.complex128, device='cuda'), lmax, zero=True)

    @staticmethod
    def from_numpy(a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        return dalm(torch.from_numpy(a).cuda(), zero=not np.any(a))

    def numpy(self):
        return self.t.cpu().numpy()

    def __len__(self):
        return self.t.numel()

    def copy(self):
        return self * 1.0

    def is_zero(self):
        return self.zero

    # ---- arithmetic
    def _lin(self, ca, other, cb):
        out = torch.empty_like(self.t)
        sht.check(sht._lib.load().plk_alm_lincomb_dev(self.t.numel(), float(ca), sht._ptr(self.t), float(cb),
                                                      sht._ptr(other.t) if other is not None else None,
                                                      sht._ptr(out), sht._stream()))
        return out

    def __add__(self, other):
        assert self.lmax == other.lmax
        return dalm(self._lin(1.0, other, 1.0), self.lmax, self.zero and other.zero)

    def __sub__(self, other):
        assert self.lmax == other.lmax
        return dalm(self._lin(1.0, other, -1.0), self.lmax, self.zero and other.zero)

    def __mul__(self, a):
        return dalm(self._lin(float(a), None, 0.0), self.lmax, self.zero or float(a) == 0.0)

    __rmul__ = __mul__

    def axpy(self, a, x):
        """self += a * x"""
        assert self.lmax == x.lmax
        sht.alm_axpy(self.t, x.t, a)
        self.zero = self.zero and x.zero
        return self

    def __iadd__(self, other):
        return self.axpy(1.0, other)

    def __isub__(self, other):
        return self.axpy(-1.0, other)

    def almxfl(self, fl_dev, inplace=False):
        out = sht.almxfl(self.t, fl_dev, out=self.t if inplace else None)
        return self if inplace else dalm(out, self.lmax, self.zero)

    def alm_copy(self, lmax=None):
        if lmax is None or lmax == self.lmax:
            return self.copy()
        assert lmax <= self.lmax
        return dalm(sht.alm_copy(self.t, lmax), lmax, self.zero)

    def alm_splice(self, alm_hi, lsplit):
        assert self.lmax >= lsplit and alm_hi.lmax >= lsplit
        return dalm(sht.alm_splice(self.t, alm_hi.t, lsplit), alm_hi.lmax, self.zero and alm_hi.zero)


class eblm:
    """(E, B) pair of alm vectors (numpy arrays or `dalm`), reference: util_alm.py:47-86."""

    def __init__(self, alm):
        elm, blm = alm
        assert len(elm) == len(blm), (len(elm), len(blm))
        self.lmax = _getlmax(len(elm))
        self.elm = elm
        self.blm = blm

    def alm_copy(self, lmax=None):
        return eblm([alm_copy(self.elm, lmax=lmax), alm_copy(self.blm, lmax=lmax)])

    def alm_splice(self, alm_hi, lsplit):
        return eblm([alm_splice(self.elm, alm_hi.elm, lsplit), alm_splice(self.blm, alm_hi.blm, lsplit)])

    def is_zero(self):
        z = [c.is_zero() if hasattr(c, 'is_zero') else not np.any(c) for c in (self.elm, self.blm)]
        return all(z)

    def copy(self):
        return self * 1.0

    def __add__(self, other):
        assert self.lmax == other.lmax
        return eblm([self.elm + other.elm, self.blm + other.blm])

    def __sub__(self, other):
        assert self.lmax == other.lmax
        return eblm([self.elm - other.elm, self.blm - other.blm])

    def __iadd__(self, other):
        assert self.lmax == other.lmax
        self.elm += other.elm
        self.blm += other.blm
        return self

    def __isub__(self, other):
        assert self.lmax == other.lmax
        self.elm -= other.elm
        self.blm -= other.blm
        return self

    def __mul__(self, other):
        return eblm([self.elm * other, self.blm * other])

    def axpy(self, a, x):
        if hasattr(self.elm, 'axpy'):
            self.elm.axpy(a, x.elm)
            self.blm.axpy(a, x.blm)
        else:
            self.elm += a * x.elm
            self.blm += a * x.blm
        return self

    def numpy(self):
        f = lambda c: c.numpy() if hasattr(c, 'numpy') else np.asarray(c)
        return f(self.elm), f(self.blm)


class teblm:
    """(T, E, B) triple of alm vectors (numpy arrays or `dalm`), reference: util_alm.py:88-160."""

    def __init__(self, alm):
        tlm, elm, blm = alm
        self.lmaxt = _getlmax(len(tlm))
        self.lmaxe = _getlmax(len(elm))
        self.lmaxb = _getlmax(len(blm))
        self.lmax = max(self.lmaxt, self.lmaxe, self.lmaxb)
        self.tlm, self.elm, self.blm = tlm, elm, blm

    def _c(self):
        return [self.tlm, self.elm, self.blm]

    def alm_copy(self, lmax=None):
        return teblm([alm_copy(c, lmax=lmax) for c in self._c()])

    def alm_splice(self, alm_hi, lsplit):
        return teblm([alm_splice(a, b, lsplit) for a, b in zip(self._c(), alm_hi._c())])

    def is_zero(self):
        return all(c.is_zero() if hasattr(c, 'is_zero') else not np.any(c) for c in self._c())

    def copy(self):
        return self * 1.0

    def _same(self, other):
        assert self.lmaxt == other.lmaxt and self.lmaxe == other.lmaxe and self.lmaxb == other.lmaxb

    def __add__(self, other):
        self._same(other)
        return teblm([a + b for a, b in zip(self._c(), other._c())])

    def __sub__(self, other):
        self._same(other)
        return teblm([a - b for a, b in zip(self._c(), other._c())])

    def __iadd__(self, other):
        self._same(other)
        self.tlm += other.tlm
        self.elm += other.elm
        self.blm += other.blm
        return self

    def __isub__(self, other):
        self._same(other)
        self.tlm -= other.tlm
        self.elm -= other.elm
        self.blm -= other.blm
        return self

    def __mul__(self, other):
        return teblm([c * other for c in self._c()])

    def axpy(self, a, x):
        for c, xc in zip(self._c(), x._c()):
            if hasattr(c, 'axpy'):
                c.axpy(a, xc)
            else:
                c += a * xc
        return self

    def numpy(self):
        f = lambda c: c.numpy() if hasattr(c, 'numpy') else np.asarray(c)
        return f(self.tlm), f(self.elm), f(self.blm)
