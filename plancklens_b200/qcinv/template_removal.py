"""Templates marginalised inside the pixel-space noise inverse (reference: plancklens/qcinv/template_removal.py).

The monopole and dipole templates -- the ones the reference's filters use by default
(filt_cinv.py:74, params/anisofilt_example.py:88-89) -- are evaluated on the GPU by libplk_b200
(`plk_map_modes_dot_dev`, `plk_map_modes_sub_dev`): mode a of (1, x, y, z) at every pixel centre, which is
what `hp.alm2map(xyz_to_alm(c))` and `alm_to_xyz(hp.map2alm(m, lmax=1)) npix/3` evaluate in the reference
(template_removal.py:134-150).  The classes here carry the bookkeeping (nmodes, which of the four modes are on).
"""
import numpy as np


class template:
    nmodes = 0
    modes = ()       # indices into (1, x, y, z)


class template_map(template):
    """One pixel-space template (reference: template_removal.py:36-53); `map` is a float64 CUDA tensor."""
    nmodes = 1

    def __init__(self, m):
        import torch
        from .. import sht
        self.map = m if isinstance(m, torch.Tensor) else sht.dev_map(m)


class template_qmap(template_map):
    """Template acting on the Q map only (reference: template_removal.py:56-82)."""
    comp = 0


class template_umap(template_map):
    """Template acting on the U map only (reference: template_removal.py:85-113)."""
    comp = 1


class template_monopole(template):
    nmodes = 1
    modes = (0,)


class template_dipole(template):
    nmodes = 3
    modes = (1, 2, 3)


def xyz_to_alm(xyz):
    """lmax = 1 alm of the map x.r (reference: template_removal.py:153-158)."""
    assert len(xyz) == 3
    alm = np.zeros(3, dtype=complex)
    alm[1] = +xyz[2] * np.sqrt(4. * np.pi / 3.)
    alm[2] = (-xyz[0] + 1.j * xyz[1]) * np.sqrt(2. * np.pi / 3.)
    return alm


def alm_to_xyz(alm):
    assert len(alm) == 3
    return np.array([-alm[2].real / np.sqrt(2. * np.pi / 3.), +alm[2].imag / np.sqrt(2. * np.pi / 3.),
                     +alm[1].real / np.sqrt(4. * np.pi / 3.)])
