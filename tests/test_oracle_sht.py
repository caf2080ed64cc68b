"""CPU: the C oracle against the brute-force definition, closed forms and adjointness."""
import numpy as np
import pytest

from helpers import alm_dot, alm_size, rand_alm, rel_l2


@pytest.mark.parametrize("spin", [0, 1, 2, 3])
def test_oracle_matches_bruteforce(oracle_sht, spin):
    from oracle import bruteforce as bf
    nside, lmax = 4, 11
    rng = np.random.default_rng(spin)
    npix = 12 * nside ** 2
    if spin == 0:
        a = rand_alm(rng, lmax)
        assert rel_l2(oracle_sht.alm2map(a, nside, lmax=lmax), bf.alm2map_spin([a], nside, 0, lmax)[0]) < 1e-13
        m = rng.standard_normal(npix)
        assert rel_l2(oracle_sht.map2alm(m, lmax=lmax), bf.map2alm_spin([m], 0, lmax)[0]) < 1e-13
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        ref = bf.alm2map_spin([g, c], nside, spin, lmax)
        got = oracle_sht.alm2map_spin([g, c], nside, spin, lmax)
        assert rel_l2(got[0], ref[0]) < 1e-13 and rel_l2(got[1], ref[1]) < 1e-13
        m = [rng.standard_normal(npix), rng.standard_normal(npix)]
        ref = bf.map2alm_spin(m, spin, lmax)
        got = oracle_sht.map2alm_spin(m, spin, lmax=lmax)
        assert rel_l2(got[0], ref[0]) < 1e-13 and rel_l2(got[1], ref[1]) < 1e-13


def test_spin_harmonic_closed_forms():
    """1Y_11 = -sqrt(3/16pi)(1 - cos) and 2Y_22 = 1/8 sqrt(5/pi)(1 - cos)^2."""
    from oracle.bruteforce import slam
    th = np.linspace(0.05, 3.0, 17)
    assert np.max(np.abs(slam(1, 1, 1, th) + np.sqrt(3 / (16 * np.pi)) * (1 - np.cos(th)))) < 1e-15
    assert np.max(np.abs(slam(2, 2, 2, th) - np.sqrt(5 / np.pi) / 8 * (1 - np.cos(th)) ** 2)) < 1e-15


@pytest.mark.parametrize("spin", [0, 2])
def test_oracle_adjointness(oracle_sht, spin):
    nside, lmax = 32, 64
    rng = np.random.default_rng(5)
    npix = 12 * nside ** 2
    w = 4 * np.pi / npix
    if spin == 0:
        a, m = rand_alm(rng, lmax), rng.standard_normal(npix)
        lhs = alm_dot(oracle_sht.map2alm(m, lmax=lmax), a, lmax)
        rhs = w * np.dot(m, oracle_sht.alm2map(a, nside, lmax=lmax))
    else:
        g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
        m = [rng.standard_normal(npix), rng.standard_normal(npix)]
        gm, cm = oracle_sht.map2alm_spin(m, spin, lmax=lmax)
        y = oracle_sht.alm2map_spin([g, c], nside, spin, lmax)
        lhs = alm_dot(gm, g, lmax) + alm_dot(cm, c, lmax)
        rhs = w * (np.dot(m[0], y[0]) + np.dot(m[1], y[1]))
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


def test_oracle_deep_underflow_against_mpmath(oracle_sht):
    """lambda_lm at l = m = 300 on a polar ring (value ~1e-470 at the pole side, ~1 at the equator)."""
    import mpmath as mp
    mp.mp.dps = 40
    nside, lmax, l, m = 128, 300, 300, 300
    a = np.zeros(alm_size(lmax), dtype=complex)
    a[m * (2 * lmax + 1 - m) // 2 + l] = 1.0
    X, _ = oracle_sht.legendre_synth(nside, 0, lmax, lmax, a)
    for r in (60, 127, 255):
        i = r + 1
        z = mp.mpf(1) - mp.mpf(i * i) / (3 * nside * nside) if i < nside else mp.mpf(2 * (2 * nside - i)) / (3 * nside)
        s = mp.sqrt(1 - z * z)
        v = mp.sqrt(1 / (4 * mp.pi))
        for k in range(1, m + 1):
            v *= -mp.sqrt(mp.mpf(2 * k + 1) / (2 * k)) * s
        if abs(v) > mp.mpf(10) ** -250:
            assert abs(X[r, m].real - float(v)) < 1e-13 * abs(float(v))


def test_spin0_against_scipy_spherical_harmonics(oracle_sht):
    """Third-party known answer for the scalar seam: scipy.special.sph_harm_y (Condon-Shortley phase, the convention
    healpy shares) against the oracle's lambda_lm e^{i m phi}, and a full spin-0 synthesis at nside 4 evaluated pixel
    by pixel as sum_lm a_lm Y_lm(theta_p, phi_p) with a_{l,-m} = (-1)^m conj(a_lm)."""
    from scipy.special import sph_harm_y
    from oracle import ref_geom as rg
    from oracle.bruteforce import slam
    rng = np.random.default_rng(0)
    for _ in range(100):
        l = int(rng.integers(0, 40))
        m = int(rng.integers(0, l + 1))
        th, ph = rng.uniform(0.01, 3.13), rng.uniform(0, 6.28)
        ref = sph_harm_y(l, m, th, ph)
        got = slam(0, l, m, np.array([th]))[0] * np.exp(1j * m * ph)
        assert abs(ref - got) <= 1e-12 * max(abs(ref), 1e-3), (l, m)
    nside, lmax = 4, 9
    theta, phi = rg.pix2ang(nside)
    a = rand_alm(rng, lmax)
    ref = np.zeros(12 * nside ** 2)
    for l in range(lmax + 1):
        for m in range(l + 1):
            c = a[rg.alm_getidx(lmax, l, m)] * sph_harm_y(l, m, theta, phi)
            ref += c.real if m == 0 else 2 * c.real
    assert rel_l2(oracle_sht.alm2map(a, nside, lmax=lmax), ref) < 1e-13


def test_healpix_ring_geometry_known_answers():
    """RING pixel centres against the published HEALPix definition (Gorski et al. 2005, eqs. 2-9; ring pixels start at
    phi = 0 on unshifted rings, as the twelve base-pixel centres of nside 1 fix: 0-3 at z = 2/3, phi = pi/4 + k pi/2;
    4-7 on the equator at phi = k pi/2; 8-11 at z = -2/3) written out by hand for nside 1 and 2, plus the structural
    invariants at nside 64: equal-area rings (z spacing), north / south mirror symmetry, ring lengths."""
    from oracle import ref_geom as rg
    th, ph = rg.pix2ang(1)
    z = np.cos(th)
    assert np.allclose(z, [2 / 3.] * 4 + [0.] * 4 + [-2 / 3.] * 4, atol=1e-15)
    assert np.allclose(ph[:4], np.pi / 4 + np.pi / 2 * np.arange(4)) and np.allclose(ph[4:8], np.pi / 2 * np.arange(4))
    assert np.allclose(ph[8:], ph[:4])
    th, ph = rg.pix2ang(2)
    z = np.cos(th)
    rings = [(4, 1 - 1 / 12.), (8, 1 - 4 / 12.), (8, 4 / 3. - 2 * 3 / 6.), (8, 0.), (8, -1 / 3.), (8, -2 / 3.), (4, -11 / 12.)]
    o = 0
    for i, (n, zr) in enumerate(rings, start=1):
        assert np.allclose(z[o:o + n], zr, atol=1e-15), i
        shifted = (i < 2 or i > 6) or (i - 2 + 1) % 2 == 1        # caps always; belt rings with (i - nside + 1) odd
        first = np.pi / n if shifted else 0.0
        assert np.allclose(ph[o:o + n], first + 2 * np.pi * np.arange(n) / n), i
        o += n
    nside = 64
    nphi, start, zr, sth, phi0 = rg.ring_info(nside)
    assert nphi.sum() == 12 * nside ** 2 and nphi.size == 4 * nside - 1
    assert np.allclose(zr, -zr[::-1], atol=1e-15) and np.array_equal(nphi, nphi[::-1])
    belt = slice(nside - 1, 3 * nside)
    assert np.allclose(np.diff(zr[belt]), -2.0 / (3 * nside)) and np.all(nphi[belt] == 4 * nside)
    i = np.arange(1, nside)
    assert np.allclose(zr[:nside - 1], 1 - i ** 2 / (3.0 * nside ** 2)) and np.array_equal(nphi[:nside - 1], 4 * i)
    assert np.allclose(sth ** 2 + zr ** 2, 1.0, atol=1e-15)
