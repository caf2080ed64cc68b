"""CPU restatement of the Wigner small-d transforms (oracle, TEST INFRASTRUCTURE ONLY).

Follows /root/reference/plancklens/wigners/wigners.f90: `wignerpos` (:566-624), `wignercoeff` (:628-684), `get_xgwg`
(:132-184), and utils_spin.wignerc (plancklens/utils_spin.py:52-93).  The Fortran cannot be compiled here (no
gfortran), so this module is pinned against the closed form  d^l_{mm'} = xi sqrt(s!(s+a+b)!/((s+a)!(s+b)!))
sin^a(theta/2) cos^b(theta/2) P_s^{(a,b)}(cos theta)  evaluated with scipy's Jacobi polynomials, and -- one level
up -- by the reference's own test identity N0 = response (tests/test_w.py), which the GPU product must satisfy.
"""
import numpy as np
from scipy.special import eval_jacobi, gammaln


def wigner_d_jacobi(l, s1, s2, x):
    """d^l_{s1 s2}(arccos x) from the Jacobi-polynomial closed form (independent of any recurrence)."""
    a, b = abs(s1 - s2), abs(s1 + s2)
    s = l - (a + b) // 2
    if s < 0:
        return np.zeros_like(x)
    xi = 1.0 if s2 >= s1 else (-1.0) ** (s2 - s1)
    lognorm = 0.5 * (gammaln(s + 1) + gammaln(s + a + b + 1) - gammaln(s + a + 1) - gammaln(s + b + 1))
    return xi * np.exp(lognorm) * (0.5 * (1 - x)) ** (0.5 * a) * (0.5 * (1 + x)) ** (0.5 * b) * eval_jacobi(s, a, b, x)


def wigner_d_all(lmax, s1, s2, x):
    """d^l_{s1 s2}(x) for all l <= lmax by upward recurrence in l, shape (lmax + 1, nx)."""
    x = np.asarray(x, dtype=float)
    l0 = max(abs(s1), abs(s2))
    out = np.zeros((lmax + 1, x.size))
    if l0 > lmax:
        return out
    out[l0] = wigner_d_jacobi(l0, s1, s2, x)
    pm, p = np.zeros_like(x), out[l0]
    m, mp = float(s1), float(s2)
    for l in range(l0, lmax):
        if l == 0:
            pn = x * p
        else:
            j = float(l)
            den = j * np.sqrt(((j + 1) ** 2 - m * m) * ((j + 1) ** 2 - mp * mp))
            pn = ((2 * j + 1) * (j * (j + 1) * x - m * mp) * p - (j + 1) * np.sqrt((j * j - m * m) * (j * j - mp * mp)) * pm) / den
        pm, p = p, pn
        out[l + 1] = p
    return out


def wignerpos(cl, x, s1, s2):
    """sum_l cl_l (2l+1)/(4 pi) d^l_{s1 s2}(x)   (wigners.f90:566-624)"""
    cl = np.asarray(cl, dtype=float)
    lmax = cl.size - 1
    d = wigner_d_all(lmax, s1, s2, x)
    return (cl * (2 * np.arange(lmax + 1) + 1) / (4 * np.pi)) @ d


def wignercoeff(f, x, s1, s2, lmax):
    """2 pi sum_x f(x) d^l_{s1 s2}(x), l <= lmax   (wigners.f90:628-684)"""
    return 2 * np.pi * (wigner_d_all(lmax, s1, s2, x) @ np.asarray(f, dtype=float))


def get_xgwg(n):
    """Gauss-Legendre nodes and weights on [-1, 1] by Newton iterations on P_n (wigners.f90:132-184)."""
    k = np.arange(1, n + 1)
    z = np.cos(np.pi * (k - 0.25) / (n + 0.5))
    for _ in range(100):
        p0, p1 = np.ones_like(z), z.copy()
        for j in range(1, n):
            p0, p1 = p1, ((2 * j + 1) * z * p1 - j * p0) / (j + 1)
        pp = n * (z * p1 - p0) / (z * z - 1)
        dz = p1 / pp
        z = z - dz
        if np.max(np.abs(dz)) < 1e-15:
            break
    p0, p1 = np.ones_like(z), z.copy()
    for j in range(1, n):
        p0, p1 = p1, ((2 * j + 1) * z * p1 - j * p0) / (j + 1)
    pp = n * (z * p1 - p0) / (z * z - 1)
    w = 2.0 / ((1 - z * z) * pp * pp)
    return z[::-1].copy(), w[::-1].copy()


def wignerc(cl1, cl2, sp1, s1, sp2, s2, lmax_out=None):
    """Legendre coefficients of xi_{sp1 s1} xi_{sp2 s2}   (utils_spin.py:52-93)"""
    lmax1, lmax2 = len(cl1) - 1, len(cl2) - 1
    lmax_out = lmax1 + lmax2 if lmax_out is None else lmax_out
    lmaxtot = lmax1 + lmax2 + lmax_out
    if not (np.any(cl1) and np.any(cl2)):
        return np.zeros(lmax_out + 1)
    n = (lmaxtot + 2 - lmaxtot % 2) // 2
    xg, wg = get_xgwg(n)
    pos = lambda cl, a, b: wignerpos(np.real(cl), xg, a, b) + (1j * wignerpos(np.imag(cl), xg, a, b) if np.iscomplexobj(cl) else 0)
    f = pos(cl1, sp1, s1) * pos(cl2, sp2, s2) * wg
    if np.iscomplexobj(f):
        return wignercoeff(f.real, xg, sp1 + sp2, s1 + s2, lmax_out) + 1j * wignercoeff(f.imag, xg, sp1 + sp2, s1 + s2, lmax_out)
    return wignercoeff(f, xg, sp1 + sp2, s1 + s2, lmax_out)
