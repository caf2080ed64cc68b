"""Wigner small-d transforms on the GPU with the call signatures of the reference's Fortran extension
(`plancklens.wigners.wigners`, wigners/wigners.f90; f2py argument order as used in utils_spin.py:52-93).

    get_xgwg(x1, x2, n)              Gauss-Legendre nodes / weights (host, cached by the caller)
    wignerpos(cl, x, s1, s2)         sum_l cl_l (2l+1)/(4 pi) d^l_{s1 s2}(x)
    wignercoeff(xiw, x, s1, s2, lmax) 2 pi sum_x xiw(x) d^l_{s1 s2}(x), l <= lmax

numpy in, numpy out; the recurrences run in libplk_b200 (`plk_wignerpos_dev`, `plk_wignercoeff_dev`).
"""
import numpy as np
import torch

from . import sht
from ._lib import check, load


def get_xgwg(x1, x2, n):
    """Gauss-Legendre abscissas and weights on [x1, x2] (reference: wigners.f90:132-184): Newton iterations on P_n
    from the Chebyshev-like first guess, vectorised over the roots."""
    n = int(n)
    k = np.arange(1, n + 1)
    z = np.cos(np.pi * (k - 0.25) / (n + 0.5))

    def legp(z):
        p0, p1 = np.ones_like(z), z.copy()
        for j in range(1, n):
            p0, p1 = p1, ((2 * j + 1) * z * p1 - j * p0) / (j + 1)
        return p1, n * (z * p1 - p0) / (z * z - 1)
    for _ in range(100):
        p, pp = legp(z)
        dz = p / pp
        z = z - dz
        if np.max(np.abs(dz)) < 1e-15:
            break
    p, pp = legp(z)
    w = 2.0 / ((1 - z * z) * pp * pp)
    xm, xl = 0.5 * (x2 + x1), 0.5 * (x2 - x1)
    return (xm + xl * z[::-1]).copy(), (xl * w[::-1]).copy()


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def wignerpos(cl, x, s1, s2):
    cl = np.asarray(cl, dtype=float)
    xd = x if isinstance(x, torch.Tensor) else _d(x)
    out = torch.empty(xd.numel(), dtype=torch.float64, device='cuda')
    check(load().plk_wignerpos_dev(sht._ptr(_d(cl)), cl.size - 1, sht._ptr(xd), int(xd.numel()), int(s1), int(s2),
                                   sht._ptr(out), sht._stream()))
    return out.cpu().numpy()


def wignercoeff(xiw, x, s1, s2, lmax):
    xd = x if isinstance(x, torch.Tensor) else _d(x)
    out = torch.empty(lmax + 1, dtype=torch.float64, device='cuda')
    check(load().plk_wignercoeff_dev(sht._ptr(_d(xiw)), sht._ptr(xd), int(xd.numel()), int(s1), int(s2), int(lmax),
                                     sht._ptr(out), sht._stream()))
    return out.cpu().numpy()
