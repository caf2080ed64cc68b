"""Brute-force spin-weighted SHT from the textbook definition (oracle, test infrastructure only).

Independent of the recurrences used anywhere else in this repository: the Wigner small-d
functions come from the Jacobi-polynomial closed form (scipy `eval_jacobi`) and every map
value is an explicit sum over (l, m).  Usable up to lmax ~ 150; meant for nside <= 32.

Conventions (`plancklens/utils_spin.py:1-14` of the reference):
  sY_lm(theta,phi) = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta) e^{i m phi}
  +|s| a_lm = -(G_lm + i C_lm),   -|s| a_lm = -(-1)^s (G_lm - i C_lm)
  spin-s map pair = (Re, Im) of sum_lm  +s a_lm  +sY_lm,
  spin-0 (healpy scalar convention): T = sum_lm a_lm Y_lm with a_{l,-m} = (-1)^m conj(a_lm).
Analysis is the exact adjoint times 4 pi / npix (healpy `map2alm(iter=0)`, uniform weights).
"""
import numpy as np
from scipy.special import eval_jacobi, gammaln

from . import ref_geom as rg


def wigner_d(l, m1, m2, theta):
    """d^l_{m1,m2}(theta) via Jacobi polynomials (Varshalovich 4.3.4), any sign of m1, m2."""
    theta = np.asarray(theta, dtype=float)
    if abs(m1) > l or abs(m2) > l:
        return np.zeros_like(theta)
    # reduce to the case  mu = m1 - m2 >= 0 , nu = m1 + m2 >= 0  with symmetries
    #   d_{m1 m2} = (-1)^{m1-m2} d_{m2 m1} = d_{-m2,-m1}
    sgn = 1.0
    a, b = m1, m2
    if a - b < 0:           # swap
        sgn *= (-1.0) ** (a - b)
        a, b = b, a
    if a + b < 0:           # (a,b) -> (-b,-a)
        a, b = -b, -a
    mu, nu = a - b, a + b   # both >= 0 now
    k = l - a
    lognorm = 0.5 * (gammaln(l + a + 1) + gammaln(l - a + 1) - gammaln(l + b + 1) - gammaln(l - b + 1))
    # d^l_{ab}(theta) = sqrt[(l+a)!(l-a)!/((l+b)!(l-b)!)] (-1)^{a-b}... sign fixed below
    s2, c2 = np.sin(theta / 2.0), np.cos(theta / 2.0)
    val = np.exp(lognorm) * s2 ** mu * c2 ** nu * eval_jacobi(k, mu, nu, np.cos(theta))
    # closed form holds for d^l_{ab} with a>=b: sign (-1)^{a-b}  (checked against mpmath in tests)
    return sgn * (-1.0) ** mu * val


def slam(s, l, m, theta):
    """sLambda_lm(theta):  sY_lm = slam * e^{i m phi}."""
    return (-1.0) ** m * np.sqrt((2 * l + 1) / (4.0 * np.pi)) * wigner_d(l, -m, s, theta)


def _ring_tables(nside):
    nphi, start, z, sth, phi0 = rg.ring_info(nside)
    theta = np.arctan2(sth, z)
    return nphi, start, theta, phi0


def alm2map_spin(gclm, nside, spin, lmax):
    """Brute-force spin-s synthesis, spin >= 0.  spin 0: gclm = [alm] -> [T]."""
    nphi, start, theta, phi0 = _ring_tables(nside)
    npix = 12 * nside * nside
    if spin == 0:
        out = np.zeros(npix)
        alm = np.asarray(gclm[0])
        for m in range(lmax + 1):
            Fm = np.zeros(theta.size, dtype=complex)
            for l in range(m, lmax + 1):
                Fm += alm[rg.alm_getidx(lmax, l, m)] * slam(0, l, m, theta)
            for k in range(theta.size):
                phi = phi0[k] + 2 * np.pi * np.arange(nphi[k]) / nphi[k]
                c = Fm[k] * np.exp(1j * m * phi)
                out[start[k]:start[k] + nphi[k]] += c.real if m == 0 else 2 * c.real
        return [out]
    G, C = np.asarray(gclm[0]), np.asarray(gclm[1])
    S = np.zeros(npix, dtype=complex)
    for m in range(lmax + 1):
        Ap = np.zeros(theta.size, dtype=complex)
        Am = np.zeros(theta.size, dtype=complex)
        for l in range(max(m, spin), lmax + 1):
            i = rg.alm_getidx(lmax, l, m)
            ap = -(G[i] + 1j * C[i])
            am = -((-1.0) ** spin) * (G[i] - 1j * C[i])
            Ap += ap * slam(spin, l, m, theta)
            Am += am * slam(-spin, l, m, theta)
        for k in range(theta.size):
            phi = phi0[k] + 2 * np.pi * np.arange(nphi[k]) / nphi[k]
            e = np.exp(1j * m * phi)
            sl = slice(start[k], start[k] + nphi[k])
            # +s field: m>=0 term, plus the m<0 term  conj(-s a_lm -sY_lm)  (a_{l,-m} symmetry of G, C)
            S[sl] += Ap[k] * e
            if m > 0:
                S[sl] += np.conj(Am[k] * e)
    return [S.real.copy(), S.imag.copy()]


def map2alm_spin(maps, spin, lmax):
    """Brute-force adjoint (times 4pi/npix).  spin 0: maps=[T] -> [alm]."""
    npix = len(maps[0])
    nside = rg.npix2nside(npix)
    nphi, start, theta, phi0 = _ring_tables(nside)
    w = 4.0 * np.pi / npix
    nalm = rg.alm_getsize(lmax)
    if spin == 0:
        alm = np.zeros(nalm, dtype=complex)
        T = np.asarray(maps[0])
        for m in range(lmax + 1):
            Fm = np.zeros(theta.size, dtype=complex)
            for k in range(theta.size):
                phi = phi0[k] + 2 * np.pi * np.arange(nphi[k]) / nphi[k]
                Fm[k] = np.sum(T[start[k]:start[k] + nphi[k]] * np.exp(-1j * m * phi))
            for l in range(m, lmax + 1):
                alm[rg.alm_getidx(lmax, l, m)] = w * np.sum(Fm * slam(0, l, m, theta))
        return [alm]
    P = np.asarray(maps[0]) + 1j * np.asarray(maps[1])      # +s field
    Pc = np.conj(P)                                           # -s field
    G = np.zeros(nalm, dtype=complex)
    C = np.zeros(nalm, dtype=complex)
    for m in range(lmax + 1):
        Fp = np.zeros(theta.size, dtype=complex)
        Fm_ = np.zeros(theta.size, dtype=complex)
        for k in range(theta.size):
            phi = phi0[k] + 2 * np.pi * np.arange(nphi[k]) / nphi[k]
            e = np.exp(-1j * m * phi)
            sl = slice(start[k], start[k] + nphi[k])
            Fp[k] = np.sum(P[sl] * e)
            Fm_[k] = np.sum(Pc[sl] * e)
        for l in range(max(m, spin), lmax + 1):
            ap = w * np.sum(Fp * slam(spin, l, m, theta))      # +s a_lm = -(G+iC)
            am = w * np.sum(Fm_ * slam(-spin, l, m, theta))    # -s a_lm = -(-1)^s (G-iC)
            am = am * (-1.0) ** spin
            i = rg.alm_getidx(lmax, l, m)
            G[i] = -0.5 * (ap + am)
            C[i] = -0.5 * (ap - am) / 1j
    return [G, C]
