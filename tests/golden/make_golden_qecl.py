"""Generates tests/golden/reference_golden_qecl.npz with the UNMODIFIED reference (see make_golden.py): qecl.library
spectra of mean-field subtracted estimates built on qest.library_sepTP.
Run from the repo root:  python tests/golden/make_golden_qecl.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'healpy_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import healpy as hp  # the shim  # noqa: E402,F401
from plancklens import qecl, qest  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

q = gi.qe_case()
out = {}
with tempfile.TemporaryDirectory() as tmp:
    iv = gi.idx_ivfs(q, hp)
    lib = qest.library_sepTP(os.path.join(tmp, 'dd'), iv, iv, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
    qcl = qecl.library(os.path.join(tmp, 'qcl'), lib, lib, np.array([1, 2, 3, 4]))
    for k1, k2 in (('ptt', 'ptt'), ('p', 'p'), ('p_p', 'ptt'), ('x', 'x')):
        out['qcl_%s_%s' % (k1, k2)] = qcl.get_sim_qcl(k1, 0, k2=k2)
    out['qcl_p_dat'] = qcl.get_sim_qcl('p', -1)
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_qecl.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
