"""Generates tests/golden/reference_golden_qest_keys.npz with the UNMODIFIED reference qest.library_sepTP (set-up as in
make_golden.py): the estimator keys beyond the lensing fast path -- point-source, noise-inhomogeneity, modulation and
polarization-rotation estimators, single-pair lensing keys ('pte', 'peb', ...), their symmetric sums, and bias
hardening -- for identical legs (dd) and for two different filtering libraries (ds, symmetrised).
Run from the repo root:  python tests/golden/make_golden_qest_keys.py
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

import healpy as hp  # the shim  # noqa: E402
from plancklens import qest  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

q = gi.qe_case()
out = {}
with tempfile.TemporaryDirectory() as tmp:
    iv1, iv2 = gi.idx_ivfs(q, hp), gi.idx_ivfs(q, hp, shift=3)
    resp = gi.toy_resplib(q['lmax_qlm'])
    dd = qest.library_sepTP(os.path.join(tmp, 'dd'), iv1, iv1, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'], resplib=resp)
    ds = qest.library_sepTP(os.path.join(tmp, 'ds'), iv1, iv2, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'], resplib=resp)
    for k in gi.QEST_EXTRA_KEYS:
        out['dd_' + k] = dd.get_sim_qlm(k, 1)
        out['ds_' + k] = ds.get_sim_qlm(k, 1)
        print(k, np.linalg.norm(out['dd_' + k]), np.linalg.norm(out['ds_' + k]))
    out['dd_mf_p_eb'] = dd.get_sim_qlm_mf('p_eb', [0, 1])
    out['dd_mf_ptt_bh_s'] = dd.get_sim_qlm_mf('ptt_bh_s', [0, 1])
    out['fundkeys'] = np.array(dd.get_fundkeys(['p_tp', 'ptt_bh_s', 'p_eb', 'stt', 'x_te']))
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_qest_keys.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, len(out), 'arrays')
