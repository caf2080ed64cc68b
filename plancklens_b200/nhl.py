"""Semi-analytical Gaussian noise levels (N0) of the quadratic estimators (reference: plancklens/nhl.py:15-96).

`get_nhl` is a sum of Wigner-d correlation functions of the estimator weights and the filtered-map spectra; the
transforms run on the GPU (`utils_spin.wignerc` -> libplk_b200) where the reference calls its Fortran extension.
The simulation-based `nhl_lib_simple` (sqlite cache around the same function) is not mirrored.
"""
import numpy as np

from . import qresp, utils
from . import utils_spin as uspin


def get_nhl(qe_key1, qe_key2, cls_weights, cls_ivfs, lmax_ivf1, lmax_ivf2, lmax_out=None, lmax_ivf12=None,
            lmax_ivf22=None, cls_weights2=None, cls_ivfs_bb=None, cls_ivfs_ab=None, cls_ivfs_ba=None):
    """(Semi-)analytical noise level of the cross-spectrum of two QE keys: (GG, CC, GC, CG) (reference: nhl.py:15-43)."""
    if lmax_ivf12 is None:
        lmax_ivf12 = lmax_ivf1
    if lmax_ivf22 is None:
        lmax_ivf22 = lmax_ivf2
    if cls_weights2 is None:
        cls_weights2 = cls_weights
    qes1 = qresp.get_qes(qe_key1, lmax_ivf1, cls_weights, lmax2=lmax_ivf12)
    qes2 = qresp.get_qes(qe_key2, lmax_ivf2, cls_weights2, lmax2=lmax_ivf22)
    if lmax_out is None:
        lmax_out = max(lmax_ivf1, lmax_ivf12) + max(lmax_ivf2, lmax_ivf22)
    return _get_nhl(qes1, qes2, cls_ivfs, lmax_out, cls_ivfs_bb=cls_ivfs_bb, cls_ivfs_ab=cls_ivfs_ab, cls_ivfs_ba=cls_ivfs_ba)


def _get_nhl(qes1, qes2, cls_ivfs, lmax_out, cls_ivfs_bb=None, cls_ivfs_ab=None, cls_ivfs_ba=None):
    """reference: nhl.py:45-96 (same loops, same order of accumulation)"""
    GG_N0 = np.zeros(lmax_out + 1, dtype=float)
    CC_N0 = np.zeros(lmax_out + 1, dtype=float)
    GC_N0 = np.zeros(lmax_out + 1, dtype=float)
    CG_N0 = np.zeros(lmax_out + 1, dtype=float)
    cls_ivfs_aa = cls_ivfs
    cls_ivfs_bb = cls_ivfs if cls_ivfs_bb is None else cls_ivfs_bb
    cls_ivfs_ab = cls_ivfs if cls_ivfs_ab is None else cls_ivfs_ab
    cls_ivfs_ba = cls_ivfs if cls_ivfs_ba is None else cls_ivfs_ba
    j, w, sc = utils.joincls, uspin.wignerc, uspin.spin_cls
    Ls = np.arange(lmax_out + 1)
    for qe1 in qes1:
        cL1 = qe1.cL(Ls)
        for qe2 in qes2:
            cL2 = qe2.cL(Ls)
            si, ti, ui, vi = (qe1.leg_a.spin_in, qe1.leg_b.spin_in, qe2.leg_a.spin_in, qe2.leg_b.spin_in)
            so, to, uo, vo = (qe1.leg_a.spin_ou, qe1.leg_b.spin_ou, qe2.leg_a.spin_ou, qe2.leg_b.spin_ou)
            assert so + to >= 0 and uo + vo >= 0, (so, to, uo, vo)
            a1, b1, a2, b2 = qe1.leg_a.cl, qe1.leg_b.cl, qe2.leg_a.cl, qe2.leg_b.cl

            clsu = j([a1, a2.conj(), sc(si, ui, cls_ivfs_aa)])
            cltv = j([b1, b2.conj(), sc(ti, vi, cls_ivfs_bb)])
            R_sutv = j([w(clsu, cltv, so, uo, to, vo, lmax_out=lmax_out), cL1, cL2])
            clsv = j([a1, b2.conj(), sc(si, vi, cls_ivfs_ab)])
            cltu = j([b1, a2.conj(), sc(ti, ui, cls_ivfs_ba)])
            R_sutv = R_sutv + j([w(clsv, cltu, so, vo, to, uo, lmax_out=lmax_out), cL1, cL2])

            # -s -t u v
            sgnms = (-1) ** (si + so)
            sgnmt = (-1) ** (ti + to)
            clsu = j([sgnms * a1.conj(), a2.conj(), sc(-si, ui, cls_ivfs_aa)])
            cltv = j([sgnmt * b1.conj(), b2.conj(), sc(-ti, vi, cls_ivfs_bb)])
            R_msmtuv = j([w(clsu, cltv, -so, uo, -to, vo, lmax_out=lmax_out), cL1, cL2])
            clsv = j([sgnms * a1.conj(), b2.conj(), sc(-si, vi, cls_ivfs_ab)])
            cltu = j([sgnmt * b1.conj(), a2.conj(), sc(-ti, ui, cls_ivfs_ba)])
            R_msmtuv = R_msmtuv + j([w(clsv, cltu, -so, vo, -to, uo, lmax_out=lmax_out), cL1, cL2])

            sg = (-1) ** (to + so)
            GG_N0 += 0.5 * R_sutv.real + 0.5 * sg * R_msmtuv.real
            CC_N0 += 0.5 * R_sutv.real - 0.5 * sg * R_msmtuv.real
            GC_N0 -= 0.5 * R_sutv.imag + 0.5 * sg * R_msmtuv.imag
            CG_N0 += 0.5 * R_sutv.imag - 0.5 * sg * R_msmtuv.imag
    return GG_N0, CC_N0, GC_N0, CG_N0
