r"""Convergence-map Wiener / inverse-variance filter operators on the GPU (reference: plancklens/qcinv/opfilt_kk.py).

 :math:`S^{-1} (S^{-1} + Y^t N^{-1} Y)^{-1} Y^t N^{-1}` with the signal spectrum
 :math:`C_L^{\kappa\kappa} = (L (L + 1) / 2)^2 C_L^{\phi\phi}` built from `s_cls['pp']`.

Apart from where the signal spectrum comes from, the operators are those of the temperature filter (the reference
module is a copy of opfilt_tt with `cltt` replaced, opfilt_kk.py:35-101); here they are the `opfilt_tt` classes fed
with the kappa spectrum, so every kernel, the template projection and the CUDA-graph path are shared.
"""
import numpy as np

from ..utils import clhash
from . import dense
from . import opfilt_tt as _tt


def p2k(lmax):
    return 0.5 * np.arange(lmax + 1) * np.arange(1, lmax + 2, dtype=float)


def pp2kk(lmax):
    return p2k(lmax) ** 2


def _as_tt(s_cls):
    """{'pp': C^phiphi} -> the dictionary opfilt_tt expects, holding C^kappakappa"""
    clpp = np.asarray(s_cls['pp'], dtype=float)
    return {'tt': clpp * pp2kk(len(clpp) - 1)}


def calc_prep(m, s_cls, n_inv_filt):
    """b = B^t N^{-1} d  (reference: opfilt_kk.py:35-41)."""
    return _tt.calc_prep(m, None, n_inv_filt)


def apply_fini(alm, s_cls, n_inv_filt):
    """Wiener-filtered klm -> inverse-variance filtered klm, in place (reference: opfilt_kk.py:43-45)."""
    _tt.apply_fini(alm, _as_tt(s_cls), n_inv_filt)


dot_op = _tt.dot_op


class fwd_op(_tt.fwd_op):
    """A x = C_L^{kk,-1} x + B^t N^{-1} B x  (reference: opfilt_kk.py:60-79)."""

    def __init__(self, s_cls, n_inv_filt):
        super().__init__(_as_tt(s_cls), n_inv_filt)
        self.clkk_inv = self.cltt_inv

    def hashdict(self):
        return {'clkk_inv': clhash(self.clkk_inv), 'n_inv_filt': self.n_inv_filt.hashdict()}


class pre_op_diag(_tt.pre_op_diag):
    """Harmonic-space diagonal preconditioner (reference: opfilt_kk.py:82-99)."""

    def __init__(self, s_cls, n_inv_filt):
        super().__init__(_as_tt(s_cls), n_inv_filt)


def pre_op_dense(lmax, fwd_op, cache_fname=None):
    return dense.pre_op_dense_kk(lmax, fwd_op, cache_fname=cache_fname)


class alm_filter_ninv(_tt.alm_filter_ninv):
    """Pixel-space inverse noise of the kappa map with monopole / dipole / template-map marginalisation
    (reference: opfilt_kk.py:105-210)."""

    def __init__(self, n_inv, b_transf, marge_monopole=False, marge_dipole=False, marge_uptolmin=-1, marge_maps=(),
                 nlev_fkl=None):
        super().__init__(n_inv, b_transf, marge_monopole=marge_monopole, marge_dipole=marge_dipole,
                         marge_uptolmin=marge_uptolmin, marge_maps=marge_maps, nlev_ftl=nlev_fkl)
        self.nlev_fkl = self.nlev_ftl

    def get_fkl(self):
        return self.get_ftl()

    def degrade(self, nside):
        if nside == self.nside:
            return self
        from .. import hp
        print("DEGRADING WITH NO MARGE MAPS")
        return alm_filter_ninv(hp.ud_grade(self.n_inv, nside, power=-2), self.b_transf,
                               marge_monopole=self.marge_monopole, marge_dipole=self.marge_dipole,
                               marge_uptolmin=self.marge_uptolmin, marge_maps=[])
