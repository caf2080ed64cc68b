"""Generates tests/golden/reference_golden_n0s.npz with the UNMODIFIED reference n0s.get_N0 (set-up as in
make_golden_resp.py: healpy shim, Wigner transforms served by the CPU oracle).

The reference's get_N0 reads an undefined name `cls_glen` (n0s.py:190); it is provided here as a module global equal
to the `cls_len` argument, which is what the docstring describes.
Run from the repo root:  python tests/golden/make_golden_n0s.py
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
from plancklens import n0s, utils, utils_spin  # noqa: E402  (reference)
from oracle import ref_wigner  # noqa: E402

import golden_inputs as gi  # noqa: E402

utils_spin.HASWIGNER = True
utils_spin.wignerc = ref_wigner.wignerc

out = {}
for name, kw in gi.n0s_cases().items():
    n0s.cls_glen = kw['cls_len']
    N0, N0c = n0s.get_N0(**kw)
    for k in N0:
        out['%s_G_%s' % (name, k)] = N0[k]
        out['%s_C_%s' % (name, k)] = N0c[k]
    print(name, sorted(N0.keys()))
a = {'tt': np.arange(5.) + 1, 'ee': np.arange(7.) + 2, 'te': 0.1 * np.arange(6.), 'bb': 0.5 * np.ones(4)}
b = {'tt': np.ones(7), 'ee': 2 * np.ones(7), 'bb': 3 * np.ones(7), 'tb': 0.2 * np.ones(7)}
out['clsdot_ab'] = utils.cls_dot([a, b])
out['clsdot_aba'] = utils.cls_dot([a, b, a])
d = utils.cls_dot([a, b, a], ret_dict=True)
for k, v in d.items():
    out['clsdot_aba_' + k] = v
dls, cldd = n0s.cls2dls({'tt': a['tt'], 'te': a['te'], 'pp': np.arange(8.) * 1e-3})
out['cls2dls'] = dls
out['cls2dls_dd'] = cldd
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_n0s.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, len(out), 'arrays')
