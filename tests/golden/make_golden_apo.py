"""Generates tests/golden/reference_golden_apo.npz: the UNMODIFIED reference utils.apodize_mask ('gaussian' and 'hybrid')
on a small cut-sky mask, with hp.smoothing served by the healpy shim (map2alm with healpy's 3 refinement passes ->
Gaussian window -> alm2map, all by the CPU oracle); plus a bare map2alm(iter=3) of a seeded map for the iteration test.
Run from the repo root:  python tests/golden/make_golden_apo.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'healpy_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import healpy as hp  # the shim  # noqa: E402
from plancklens import utils  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

a = gi.apo_case()
out = {'apo_gaussian': utils.apodize_mask(a['mask'], sigma_arcmin=a['sigma_arcmin'], lmax=a['lmax'], method='gaussian', cache_dir=None),
       'apo_hybrid': utils.apodize_mask(a['mask'], sigma_arcmin=a['sigma_arcmin'], lmax=a['lmax'], method='hybrid', cache_dir=None),
       'alm_iter3': hp.map2alm(a['map'], lmax=a['lmax'], iter=3),
       'alm_iter1': hp.map2alm(a['map'], lmax=a['lmax'], iter=1)}
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_apo.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: v.shape for k, v in out.items()})
