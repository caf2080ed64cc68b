"""Generates tests/golden/reference_golden_n1.npz with the UNMODIFIED reference n1.library_n1 (healpy shim on the path).
The Fortran integrator `n1f` cannot be built here; a deterministic stand-in with the same argument list is attached to
the reference module (`n1.n1f`), so that everything around it -- key ordering, decomposition of derived estimators,
sampling and splining in L, both sqlite caches -- is the reference's own code.
Run from the repo root:  python tests/golden/make_golden_n1.py
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
from plancklens.n1 import n1  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

c = gi.n1_case()


class fake_n1f:
    n1l = staticmethod(gi.fake_n1l)


n1.n1f = fake_n1f
out = {}
with tempfile.TemporaryDirectory() as tmp:
    lib = n1.library_n1(os.path.join(tmp, 'n1'), c['cltt'], c['clte'], c['clee'], lmaxphi=c['lmaxphi'])
    out['lps'] = np.asarray(lib.lps)
    for kA, kB in gi.N1_PAIRS:
        out['n1_%s_%s' % (kA, kB)] = lib.get_n1(kA, 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'], kB=kB,
                                                ftlB=c['ftlB'])
    out['n1_flat'] = lib.get_n1('ptt', 'p', c['clpp'], c['ftl'], c['fel'], c['fbl'], c['Lmax'] - 10,
                                n1_flat=lambda ell: ell ** 2 * (ell + 1.) ** 2)
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_n1.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: v.shape for k, v in out.items()})
