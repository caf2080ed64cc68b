"""The SHT seam of the reference (plancklens/shts.py:4-35), served by the B200 CUDA kernels.

Same four names and argument meanings as the reference module, numpy in / numpy out.  `qcinv.opfilt_tt` and
`qcinv.opfilt_pp` of the reference import exactly these (opfilt_tt.py:16, opfilt_pp.py:16).
Device-resident callers (CG, QE) use `plancklens_b200.sht.Plan` directly and skip the host copies.
"""
import numpy as np

from . import hp

HASLENSPYX = False


def alm2map(alm, nside):
    """alm (healpy layout, complex) -> RING map at nside.  reference: shts.py:12 / :35."""
    return hp.alm2map(alm, nside)


def map2alm(m, lmax, **kwargs):
    """Analysis with uniform weights; keyword arguments go to `hp.map2alm` as in the reference (every call on its hot
    path passes iter=0: single pass).  reference: shts.py:16."""
    return hp.map2alm(m, lmax=lmax, **kwargs)


def alm2map_spin(gclm, nside, spin, lmax):
    """(G, C) alm pair -> (Re, Im) spin-s maps.  reference: shts.py:22."""
    assert len(gclm) == 2, len(gclm)
    return hp.alm2map_spin(gclm, nside, spin, lmax)


def map2alm_spin(qumap, spin, lmax):
    """(Re, Im) spin-s maps -> (G, C) alm pair.  reference: shts.py:26."""
    assert len(qumap) == 2
    assert np.size(qumap[0]) == np.size(qumap[1]), (np.size(qumap[0]), np.size(qumap[1]))
    return hp.map2alm_spin(qumap, spin, lmax=lmax)
