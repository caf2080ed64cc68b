"""Generates tests/golden/reference_golden_resp2.npz with the UNMODIFIED reference (set-up as in make_golden_resp.py,
Wigner transforms served by the CPU oracle): noise-variance-map responses (qresp._get_response_custom), response
derivatives (get_dresponse_dlncl) and the deflection-induced mean-field response (get_mf_resp).
Run from the repo root:  python tests/golden/make_golden_resp2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'healpy_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import healpy as hp  # the shim  # noqa: E402,F401
from plancklens import qresp, utils_spin  # noqa: E402  (reference)
from oracle import ref_wigner  # noqa: E402

import golden_inputs as gi  # noqa: E402

utils_spin.HASWIGNER = True
utils_spin.wignerc = ref_wigner.wignerc

r = gi.resp_case()
transf = gi.resp_transf(r['lmax'])
out = {}
for key, src in gi.RESP2_CUSTOM:
    out['custom_%s_%s' % (key, src)] = np.array(qresp.get_response(key, r['lmax'], src, r['cls_weight'], r['cls_len'], r['fal_sep'],
                                                                   lmax_qlm=r['lmax_qlm'], transf=transf))
for key, l, ck, src in gi.RESP2_DERIV:
    out['dresp_%s_%d_%s_%s' % (key, l, ck, src)] = np.array(qresp.get_dresponse_dlncl(key, l, ck, r['lmax'], src, r['cls_weight'], r['cls_len'],
                                                                                    r['fal_sep'], lmax_out=r['lmax_qlm']))
for key in ('ptt', 'p_p'):
    GL, CL, terms = qresp.get_mf_resp(key, r['cls_len'], r['cls_ivfs_sep'], r['lmax'] - 10, r['lmax_qlm'], retterms=True)
    out['mf_%s_G' % key], out['mf_%s_C' % key] = GL, CL
    for k, v in terms.items():
        out['mf_%s_%s' % (key, k)] = v
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_resp2.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, len(out), 'arrays')
