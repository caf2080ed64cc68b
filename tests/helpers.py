import numpy as np


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_ls(lmax):
    return np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])


def rand_alm(rng, lmax, lmin=0):
    """Random alm of a real field: complex Gaussian, m=0 entries real, zero below lmin."""
    n = alm_size(lmax)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a[:lmax + 1] = a[:lmax + 1].real
    a[alm_ls(lmax) < lmin] = 0
    return a


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def alm_dot(a, b, lmax):
    """sum_l sum_{m=-l..l} a_lm conj(b_lm) for alms of real fields stored for m >= 0."""
    w = np.full(a.size, 2.0)
    w[:lmax + 1] = 1.0
    return float(np.sum(w * (a * np.conj(b)).real))
