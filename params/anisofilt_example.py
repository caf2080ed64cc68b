"""Parameter file for lensing reconstruction with anisotropic (masked-sky, conjugate-gradient) filtering (B200 version).

Same structure and object names as the reference's params/anisofilt_example.py (cinv_t, cinv_p, ivfs_raw, ivfs,
qlms_dd).  The Planck lensing mask of the reference lives on NERSC; here a synthetic Galactic cut plus point-source
holes is built on the fly, and the sims are Gaussian skies with the FFP10 lensed spectra.
"""
import os

import numpy as np

import plancklens_b200
from plancklens_b200 import hp, qest, utils
from plancklens_b200.filt import filt_cinv, filt_util
from plancklens_b200.sims import cmbs, maps, phas, utils as maps_utils

assert 'PLENS' in os.environ.keys(), 'Set env. variable PLENS to a writeable folder'
TEMP = os.path.join(os.environ['PLENS'], 'temp', 'anisofilt_example')
cls_path = os.path.join(os.path.dirname(os.path.abspath(plancklens_b200.__file__)), 'data', 'cls')

nside = int(os.environ.get('PLK_NSIDE', 2048))
lmax_ivf = int(os.environ.get('PLK_LMAX_IVF', 2048))
lmin_ivf = 100
lmax_qlm = int(os.environ.get('PLK_LMAX_QLM', 4096))
nlev_t = 35.
nlev_p = 55.
nsims = int(os.environ.get('PLK_NSIMS', 300))

transf = hp.gauss_beam(5. / 60. / 180. * np.pi, lmax=lmax_ivf)
cl_len = utils.camb_clfile(os.path.join(cls_path, 'FFP10_wdipole_lensedCls.dat'), lmax=max(lmax_ivf, lmax_qlm))
cl_ivf = {k: cl_len[k][:lmax_ivf + 1] for k in ['tt', 'ee', 'bb', 'te']}

device_sims = bool(int(os.environ.get('PLK_DEVICE_SIMS', 0)))   # draw phases on the GPU (Philox kernels)
pix_phas = phas.pix_lib_phas(os.path.join(TEMP, 'pix_phas_nside%s' % nside), 3, (hp.nside2npix(nside),), device=device_sims)
cmb_sims = cmbs.sims_cmb_unl(cl_ivf, phas.lib_phas(os.path.join(TEMP, 'cmb_phas'), 3, lmax_ivf, device=device_sims))
sims = maps_utils.sim_lib_shuffle(maps.cmb_maps_nlev(cmb_sims, transf, nlev_t, nlev_p, nside, pix_lib_phas=pix_phas),
                                  {idx: nsims if idx == -1 else idx for idx in range(-1, nsims)})


def _ring_z(nside):
    z = np.empty(12 * nside ** 2)
    p = 0
    for i in range(1, 4 * nside):
        ii = i if i < nside else (nside if i <= 3 * nside else 4 * nside - i)
        zz = 1 - i * i / (3. * nside ** 2) if i < nside else ((2 * nside - i) * 2. / (3. * nside) if i <= 3 * nside else -(1 - ii * ii / (3. * nside ** 2)))
        z[p:p + 4 * ii] = zz
        p += 4 * ii
    return z


# Synthetic mask: Galactic cut |b| < 20 deg (replace with your mask)
mask = (np.abs(_ring_z(nside)) >= np.sin(np.deg2rad(20.))).astype(float)

libdir_cinvt = os.path.join(TEMP, 'cinv_t')
libdir_cinvp = os.path.join(TEMP, 'cinv_p')
libdir_ivfs = os.path.join(TEMP, 'ivfs')

# Homogeneous noise in the filter outside the masked area: inverse pixel variance = pixel area [arcmin^2] / nlev^2
vamin2 = hp.nside2pixarea(nside, degrees=True) * 3600.
ninv_t = [np.array([vamin2 / nlev_t ** 2]), mask]
cinv_t = filt_cinv.cinv_t(libdir_cinvt, lmax_ivf, nside, cl_ivf, transf, ninv_t, marge_monopole=True, marge_dipole=True, marge_maps=[])

ninv_p = [[np.array([vamin2 / nlev_p ** 2]), mask]]
cinv_p = filt_cinv.cinv_p(libdir_cinvp, lmax_ivf, nside, cl_ivf, transf, ninv_p)

ivfs_raw = filt_cinv.library_cinv_sepTP(libdir_ivfs, sims, cinv_t, cinv_p, cl_len)
ftl = np.ones(lmax_ivf + 1, dtype=float) * (np.arange(lmax_ivf + 1) >= lmin_ivf)  # rescaling or cuts. Here just a lmin cut
fel = np.ones(lmax_ivf + 1, dtype=float) * (np.arange(lmax_ivf + 1) >= lmin_ivf)
fbl = np.ones(lmax_ivf + 1, dtype=float) * (np.arange(lmax_ivf + 1) >= lmin_ivf)
ivfs = filt_util.library_ftl(ivfs_raw, lmax_ivf, ftl, fel, fbl)

qlms_dd = qest.library_sepTP(os.path.join(TEMP, 'qlms_dd'), ivfs, ivfs, cl_len['te'], nside, lmax_qlm=lmax_qlm)
