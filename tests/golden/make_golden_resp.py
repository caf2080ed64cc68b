"""Generates tests/golden/reference_golden_resp.npz with the UNMODIFIED reference host logic of qresp / nhl /
utils_spin / utils (see make_golden.py for the set-up).  The reference's Wigner transforms are a Fortran extension
that cannot be built here; `plancklens.utils_spin.wignerc` is therefore pointed at the CPU oracle
(oracle/ref_wigner.wignerc) -- the same device the healpy shim is for the SHT seam: everything ABOVE the seam
(estimator weights, spin matrices, response and N0 assembly) is the reference's own code.
Run from the repo root:  python tests/golden/make_golden_resp.py
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
from plancklens import nhl, qresp, utils, utils_spin  # noqa: E402  (reference)
from oracle import ref_wigner  # noqa: E402

import golden_inputs as gi  # noqa: E402

utils_spin.HASWIGNER = True
utils_spin.wignerc = ref_wigner.wignerc

r = gi.resp_case()
out = {}
# pure host helpers
for s1, s2 in [(0, 0), (0, 2), (2, 0), (2, 2), (2, -2), (-2, 2), (-2, 0), (0, -2)]:
    out['spinmat_%d_%d' % (s1, s2)] = np.asarray(utils_spin.get_spin_matrix(s1, s2, r['fal_tb']), dtype=complex)
    out['spincls_%d_%d' % (s1, s2)] = np.asarray(utils_spin.spin_cls(s1, s2, r['cls_ivfs_tb']), dtype=complex)
inv = utils.cl_inverse(r['cls_dat'])
for k, v in inv.items():
    out['clinv_' + k] = v
for key in ['ptt', 'p_p', 'p', 'x', 'ftt', 'pee', 'p_te', 'a_p']:
    qes = qresp.get_qes(key, r['lmax'], r['cls_weight'])
    out['qes_%s_n' % key] = np.array([len(qes)])
    for i, q in enumerate(qes):
        out['qes_%s_%d_spins' % (key, i)] = np.array([q.leg_a.spin_in, q.leg_a.spin_ou, q.leg_b.spin_in, q.leg_b.spin_ou])
        out['qes_%s_%d_cla' % (key, i)] = np.asarray(q.leg_a.cl, dtype=complex)
        out['qes_%s_%d_clb' % (key, i)] = np.asarray(q.leg_b.cl, dtype=complex)
        out['qes_%s_%d_cL' % (key, i)] = np.asarray(q.cL(np.arange(r['lmax_qlm'] + 1)), dtype=float)
# responses and N0 through the reference's assembly code (Wigner transforms by the oracle)
for key, src, fal in [('ptt', 'p', 'fal_sep'), ('p_p', 'p', 'fal_sep'), ('p', 'p', 'fal_jt'), ('x', 'x', 'fal_jt'),
                      ('ftt', 'f', 'fal_sep'), ('ptt', 'f', 'fal_sep'), ('p', 'p', 'fal_tb'), ('ptt_bh_f', 'p', 'fal_sep')]:
    R = qresp.get_response(key, r['lmax'], src, r['cls_weight'], r['cls_len'], r[fal], lmax_qlm=r['lmax_qlm'])
    out['resp_%s_%s_%s' % (key, src, fal)] = np.array(R)
for k1, k2, ivf in [('ptt', 'ptt', 'cls_ivfs_sep'), ('p_p', 'p_p', 'cls_ivfs_sep'), ('p', 'p', 'cls_ivfs_jt'),
                    ('ptt', 'p_p', 'cls_ivfs_sep'), ('p', 'p', 'cls_ivfs_tb'), ('x', 'p', 'cls_ivfs_tb')]:
    N = nhl.get_nhl(k1, k2, r['cls_weight'], r[ivf], r['lmax'], r['lmax'], lmax_out=r['lmax_qlm'])
    out['nhl_%s_%s_%s' % (k1, k2, ivf)] = np.array(N)
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_resp.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, len(out), 'arrays')
