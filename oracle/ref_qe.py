"""CPU restatement of the reference's fast quadratic-estimator path (oracle, test infrastructure only).

Follows /root/reference/plancklens/qest.py: `_get_sim_Tgclm` (:248-263), `_get_sim_Pgclm` (:265-285),
`_get_sim_MVgclm` (:318-322) and the leg builders of `lib_filt2map_sepTP` (:506-530, :566-638), for separately
filtered T and P with the C^TE cross terms.  Pinned against the unmodified reference through
tests/golden/reference_golden.npz (tests/test_oracle_golden.py).
"""
import numpy as np

from . import ref_sht as sht
from .healpy_shim.healpy import almxfl


def _lmax(alm):
    return int(np.floor(np.sqrt(2 * alm.size) - 1))


def wf_tlm(tbar, ebar, cls, k):
    """Wiener-filtered T entering the gradient leg (qest.py:579-588)."""
    t = almxfl(tbar, cls['tt'])
    if k == 'p':
        t = t + almxfl(ebar, cls['te'])
    return t


def wf_eblm(tbar, ebar, bbar, cls, k):
    """Wiener-filtered E, B entering the gradient legs (qest.py:609-618)."""
    e = almxfl(ebar, cls['ee'])
    b = almxfl(bbar, cls['bb'])
    if k == 'p':
        e = e + almxfl(tbar, cls['te'])
    return e, b


def t_products(tbar, twf, nside):
    """(G t, C t): qest.py:254-257 with get_irestmap (:506-514) and get_gtmap (:590-593)."""
    lmax = _lmax(tbar)
    l = np.arange(lmax + 1, dtype=float)
    tmap = sht.alm2map(tbar, nside, lmax=lmax)
    Glm = almxfl(twf, -np.sqrt(l * (l + 1)))
    G, C = sht.alm2map_spin([Glm, np.zeros_like(Glm)], nside, 1, lmax)
    return G * tmap, C * tmap


def p_products(ebar, bbar, ewf, bwf, nside):
    """GC = (Q - iU)(G3 + iC3) - (Q + iU)(G1 - iC1): qest.py:273-278 with get_irespmap (:521-530), get_gpmap (:620-636)."""
    lmax = _lmax(ebar)
    l = np.arange(lmax + 1, dtype=float)
    Q, U = sht.alm2map_spin([0.5 * ebar, 0.5 * bbar], nside, 2, lmax)
    GC = np.zeros(Q.size, dtype=complex)
    for spin, sgn in ((3, 1), (1, -1)):
        fl = (l - 2) * (l + 3) if spin == 3 else (l + 2) * (l - 1)
        fl[:spin] = 0.
        fl = np.sqrt(fl)
        Gs, Cs = sht.alm2map_spin([almxfl(ewf, fl), almxfl(bwf, fl)], nside, spin, lmax)
        if spin == 3:
            GC += (Q - 1j * U) * (Gs + 1j * Cs)
        else:
            GC -= (Q + 1j * U) * (Gs - 1j * Cs)
    return GC.real, GC.imag


def analyse(re, im, lmax_qlm):
    """map2alm_spin(., 1) x -sqrt(L(L+1)): qest.py:259-262, 280-284."""
    G, C = sht.map2alm_spin([re, im], 1, lmax=lmax_qlm)
    L = np.arange(lmax_qlm + 1, dtype=float)
    fl = -np.sqrt(L * (L + 1))
    return almxfl(G, fl), almxfl(C, fl)


def qe(k, tbar, ebar, bbar, cls, nside, lmax_qlm, tbar2=None, ebar2=None, bbar2=None):
    """Gradient and curl qlm for k in 'ptt', 'p_p', 'p'; leg 2 (Wiener leg) taken from the *2 alms when given."""
    t2 = tbar if tbar2 is None else tbar2
    e2 = ebar if ebar2 is None else ebar2
    b2 = bbar if bbar2 is None else bbar2
    G = C = 0
    if k in ('p_p', 'p'):
        ewf, bwf = wf_eblm(t2, e2, b2, cls, k)
        g, c = analyse(*p_products(ebar, bbar, ewf, bwf, nside), lmax_qlm)
        G, C = G + g, C + c
    if k in ('ptt', 'p'):
        twf = wf_tlm(t2, e2, cls, k)
        g, c = analyse(*t_products(tbar, twf, nside), lmax_qlm)
        G, C = G + g, C + c
    return G, C
