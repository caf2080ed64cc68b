"""Generates tests/golden/reference_golden_templates.npz with the UNMODIFIED reference (see make_golden.py for the
set-up): pixel-space template marginalisation -- opfilt_tt marge_maps (with monopole + dipole) and opfilt_pp
marge_qmaps / marge_umaps.  Run from the repo root:  python tests/golden/make_golden_templates.py
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
from plancklens.qcinv import opfilt_pp, opfilt_tt  # noqa: E402  (reference)
from plancklens.qcinv.util_alm import eblm  # noqa: E402

import golden_inputs as gi  # noqa: E402

out = {}
c = gi.cg_case()
t = gi.template_case()

for tag, kw in (('tm', dict(marge_monopole=True, marge_dipole=True, marge_maps=t['tmaps'])),
                ('tmonly', dict(marge_maps=t['tmaps'][:1]))):
    nf = opfilt_tt.alm_filter_ninv(c['ninv_t'], c['transf'], **kw)
    out[tag + '_pinv'] = nf.Pt_Nn1_P_inv
    m = c['tmap'].copy()
    nf.apply_map(m)
    out[tag + '_apply_map'] = m
    out[tag + '_fwd'] = opfilt_tt.fwd_op(c['cls'], nf)(c['x_t'].copy())
    out[tag + '_prep'] = opfilt_tt.calc_prep(c['tmap'], c['cls'], nf)

nf = opfilt_pp.alm_filter_ninv(c['ninv_p1'], c['transf'], marge_qmaps=t['qmaps'], marge_umaps=t['umaps'])
q, u = c['qmap'].copy(), c['umap'].copy()
nf.apply_map([q, u])
out['pm_apply_q'], out['pm_apply_u'] = q, u
out['pm_tniti'] = nf.tniti
r = opfilt_pp.fwd_op(c['cls'], nf)(eblm([c['x_e'].copy(), c['x_b'].copy()]))
out['pm_fwd_e'], out['pm_fwd_b'] = r.elm, r.blm

fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_templates.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
