"""CPU restatement of the device random-number generator (oracle, test infrastructure only).

Philox4x32-10 as published (Salmon, Moraes, Dror & Shaw 2011, "Parallel random numbers: as easy as 1, 2, 3"; the
Random123 library), pinned against that library's known-answer vectors in tests/test_oracle_rng.py, followed by the
uniform -> Box-Muller mapping of plancklens_b200/csrc/plk_rng.cuh.  The reference itself draws with numpy
(/root/reference/plancklens/sims/phas.py:162-168); what is restated from it is the recipe for alm phases:
(N(0,1) + i N(0,1)) / sqrt(2), real N(0,1) at m = 0.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """counter words c0..c3 (arrays of uint32 values held as uint64), key words k0, k1 (python ints) -> 4 arrays"""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)]
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        kk0 = np.uint64((k0 + r * W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ kk0
        n1 = p1 & _MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ kk1
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
    return c0, c1, c2, c3


def words(seed, stream, ncalls):
    """[ncalls, 4] uint32: call i uses counter (i lo, i hi, stream lo, stream hi) and key (seed lo, seed hi)"""
    i = np.arange(ncalls, dtype=np.uint64)
    z = np.zeros(ncalls, dtype=np.uint64)
    out = philox4x32_10(i & _MASK, i >> np.uint64(32), z + np.uint64(stream & 0xFFFFFFFF), z + np.uint64(stream >> 32),
                        seed & 0xFFFFFFFF, seed >> 32)
    return np.stack(out, 1).astype(np.uint32)


def _u53(hi, lo):
    x = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)
    return (x.astype(np.float64) + 0.5) / 9007199254740992.0


def _pairs(seed, stream, ncalls):
    w = words(seed, stream, ncalls)
    u1, u2 = _u53(w[:, 0], w[:, 1]), _u53(w[:, 2], w[:, 3])
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


def randn(seed, stream, n):
    """n unit normals: out[2 i], out[2 i + 1] from call i"""
    z0, z1 = _pairs(seed, stream, (n + 1) // 2)
    return np.stack([z0, z1], 1).reshape(-1)[:n]


def randn_alm(seed, stream, lmax):
    """alm phases of a real field (sims/phas.py:162-168): element i from call i"""
    n = (lmax + 1) * (lmax + 2) // 2
    z0, z1 = _pairs(seed, stream, n)
    a = (z0 + 1j * z1) / np.sqrt(2.)
    a[:lmax + 1] = z0[:lmax + 1]
    return a
