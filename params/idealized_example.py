"""Parameter file for lensing reconstruction on an idealized, full-sky simulation library (B200 version).

Same structure and object names as the reference's params/idealized_example.py: it instantiates
    * the inverse-variance filtered simulation library 'ivfs'
    * the three quadratic-estimator libraries 'qlms_dd', 'qlms_ds', 'qlms_ss'.
Differences forced by the environment (SURVEY.md, table of discrepancies): the FFP10 lensed CMB sims on NERSC are
replaced by Gaussian skies drawn from the FFP10 lensed spectra (`cmbs.sims_cmb_unl`), and the transfer function is
the 5' beam alone (`hp.pixwin` needs the HEALPix data files).  The spectra / response / bias libraries of the
reference file (qecl, nhl, n1, qresp) are instantiated as there; `n1_dd` serves cached N1 curves only (the flat-sky
integrator is the reference's Fortran extension, see plancklens_b200/n1/n1.py).

Sizes can be scaled down for tests through the environment: PLK_NSIDE, PLK_LMAX_IVF, PLK_LMAX_QLM, PLK_NSIMS;
PLK_DEVICE_SIMS=1 draws the simulations on the GPU.
"""
import os

import numpy as np

import plancklens_b200
from plancklens_b200 import hp, nhl, qecl, qest, qresp, utils
from plancklens_b200.filt import filt_simple, filt_util
from plancklens_b200.n1 import n1
from plancklens_b200.sims import cmbs, maps, phas, utils as maps_utils

assert 'PLENS' in os.environ.keys(), 'Set env. variable PLENS to a writeable folder'
TEMP = os.path.join(os.environ['PLENS'], 'temp', 'idealized_example')
cls_path = os.path.join(os.path.dirname(os.path.abspath(plancklens_b200.__file__)), 'data', 'cls')

# --- definition of simulation and inverse-variance filtered simulation libraries:
nside = int(os.environ.get('PLK_NSIDE', 2048))       # Healpix resolution of the data and sims.
lmax_ivf = int(os.environ.get('PLK_LMAX_IVF', 2048))
lmin_ivf = min(100, lmax_ivf // 8)  # We will use in the QE only CMB modes between lmin_ivf and lmax_ivf
lmax_qlm = int(os.environ.get('PLK_LMAX_QLM', 4096))  # We will calculate lensing estimates until multipole lmax_qlm.
nlev_t = 35.  # Filtering noise level in temperature (here also used for the noise simulations generation).
nlev_p = 55.  # Filtering noise level in polarization (here also used for the noise simulations generation).
nsims = int(os.environ.get('PLK_NSIMS', 300))  # Total number of simulations to consider.

transf = hp.gauss_beam(5. / 60. / 180. * np.pi, lmax=lmax_ivf)
#: CMB transfer function. Here a 5' Gaussian beam.

cl_len = utils.camb_clfile(os.path.join(cls_path, 'FFP10_wdipole_lensedCls.dat'), lmax=max(lmax_ivf, lmax_qlm))
#: Fiducial lensed power spectra used for the analysis.

cl_weight = utils.camb_clfile(os.path.join(cls_path, 'FFP10_wdipole_lensedCls.dat'), lmax=max(lmax_ivf, lmax_qlm))
cl_weight['bb'] *= 0.
#: CMB spectra entering the QE weights

device_sims = bool(int(os.environ.get('PLK_DEVICE_SIMS', 0)))   # draw phases on the GPU (Philox kernels)
pix_phas = phas.pix_lib_phas(os.path.join(TEMP, 'pix_phas_nside%s' % nside), 3, (hp.nside2npix(nside),), device=device_sims)
#: Noise simulation T, Q, U random phases instance.
cmb_phas = phas.lib_phas(os.path.join(TEMP, 'cmb_phas'), 3, lmax_ivf, device=device_sims)
cmb_sims = cmbs.sims_cmb_unl({k: cl_len[k][:lmax_ivf + 1] for k in ['tt', 'ee', 'bb', 'te']}, cmb_phas)
#: Gaussian CMB skies with the lensed spectra (stand-in for planck2018_sims.cmb_len_ffp10()).

sims = maps_utils.sim_lib_shuffle(maps.cmb_maps_nlev(cmb_sims, transf, nlev_t, nlev_p, nside, pix_lib_phas=pix_phas),
                                  {idx: nsims if idx == -1 else idx for idx in range(-1, nsims)})
#: Simulation library: index -1 (the "data") points at a simulation outside the analysis set.

# --- inverse-variance filtering library: trivial isotropic filtering (independent T and Pol. filtering)
ftl = utils.cli(cl_len['tt'][:lmax_ivf + 1] + (nlev_t / 60. / 180. * np.pi / transf) ** 2)
fel = utils.cli(cl_len['ee'][:lmax_ivf + 1] + (nlev_p / 60. / 180. * np.pi / transf) ** 2)
fbl = utils.cli(cl_len['bb'][:lmax_ivf + 1] + (nlev_p / 60. / 180. * np.pi / transf) ** 2)
ftl[:lmin_ivf] *= 0.
fel[:lmin_ivf] *= 0.
fbl[:lmin_ivf] *= 0.
#: Inverse CMB co-variance in T, E and B (neglecting TE coupling).

ivfs = filt_simple.library_fullsky_sepTP(os.path.join(TEMP, 'ivfs'), sims, nside, transf, cl_len, ftl, fel, fbl, cache=True)
#: Inverse-variance filtering instance.

# ---- QE libraries: same simulation on both legs (dd), simulation x data (ds), simulation x shuffled simulation (ss)
blk = max(1, min(60, nsims))
ss_dict = {k: v for k, v in zip(np.concatenate([range(i * blk, (i + 1) * blk) for i in range(0, max(1, nsims // blk))]),
                                np.concatenate([np.roll(range(i * blk, (i + 1) * blk), -1) for i in range(0, max(1, nsims // blk))]))}
ds_dict = {k: -1 for k in range(nsims)}

ivfs_d = filt_util.library_shuffle(ivfs, ds_dict)
#: This is a filtering instance always returning the data map.
ivfs_s = filt_util.library_shuffle(ivfs, ss_dict)
#: This is a filtering instance shuffling simulation indices according to 'ss_dict'.

qlms_dd = qest.library_sepTP(os.path.join(TEMP, 'qlms_dd'), ivfs, ivfs, cl_len['te'], nside, lmax_qlm=lmax_qlm)
qlms_ds = qest.library_sepTP(os.path.join(TEMP, 'qlms_ds'), ivfs, ivfs_d, cl_len['te'], nside, lmax_qlm=lmax_qlm)
qlms_ss = qest.library_sepTP(os.path.join(TEMP, 'qlms_ss'), ivfs, ivfs_s, cl_len['te'], nside, lmax_qlm=lmax_qlm)

mc_sims_bias = np.arange(min(60, nsims))  #: The mean-field will be calculated from these simulations.
mc_sims_var = np.arange(min(60, nsims), nsims)  #: The covariance matrix will be calculated from these simulations

# ---- QE spectra libraries: power spectra of the QE maps after mean-field subtraction (only qcls_dd needs one)
mc_sims_mf_dd = mc_sims_bias
mc_sims_mf_ds = np.array([], dtype=int)
mc_sims_mf_ss = np.array([], dtype=int)

qcls_dd = qecl.library(os.path.join(TEMP, 'qcls_dd'), qlms_dd, qlms_dd, mc_sims_mf_dd)
qcls_ds = qecl.library(os.path.join(TEMP, 'qcls_ds'), qlms_ds, qlms_ds, mc_sims_mf_ds)
qcls_ss = qecl.library(os.path.join(TEMP, 'qcls_ss'), qlms_ss, qlms_ss, mc_sims_mf_ss)

# ---- semi-analytical Gaussian lensing bias library
nhl_dd = nhl.nhl_lib_simple(os.path.join(TEMP, 'nhl_dd'), ivfs, cl_weight, lmax_qlm)

# ---- N1 lensing bias library (constructor, hash and sqlite caches as in the reference; no integrator in this package)
libdir_n1_dd = os.path.join(TEMP, 'n1_ffp10')
n1_dd = n1.library_n1(libdir_n1_dd, cl_len['tt'], cl_len['te'], cl_len['ee'])

# ---- QE response calculation library
qresp_dd = qresp.resp_lib_simple(os.path.join(TEMP, 'qresp'), lmax_ivf, cl_weight, cl_len,
                                 {'t': ivfs.get_ftl(), 'e': ivfs.get_fel(), 'b': ivfs.get_fbl()}, lmax_qlm)
