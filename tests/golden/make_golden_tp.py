"""Generates tests/golden/reference_golden_tp.npz with the UNMODIFIED reference (see make_golden.py for the set-up):
joint temperature + polarization filter, qcinv/opfilt_tp.py -- operators and a two-level multigrid CG solve.
Run from the repo root:  python tests/golden/make_golden_tp.py
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
from plancklens.qcinv import cd_solve, multigrid, opfilt_tp  # noqa: E402  (reference)
from plancklens.qcinv.util_alm import teblm  # noqa: E402

import golden_inputs as gi  # noqa: E402

out = {}
c = gi.cg_case()
t = gi.template_case()
n = hp.Alm.getsize(c['lmax'])
x = lambda: teblm([c['x_t'].copy(), c['x_e'].copy(), c['x_b'].copy()])

for tag, ninv, kw in (('tp2', [c['ninv_t'][0], c['ninv_p1'][0][0]], dict(marge_monopole=True, marge_dipole=True)),
                      ('tp4', [c['ninv_t'][0]] + [m[0] for m in c['ninv_p3']], dict(marge_maps_t=t['tmaps'][:1]))):
    nf = opfilt_tp.alm_filter_ninv(ninv, c['transf'], **kw)
    fwd = opfilt_tp.fwd_op(c['cls'], nf)
    r = fwd(x())
    out[tag + '_fwd_t'], out[tag + '_fwd_e'], out[tag + '_fwd_b'] = r.tlm, r.elm, r.blm
    p = opfilt_tp.calc_prep([c['tmap'], c['qmap'], c['umap']], c['cls'], nf)
    out[tag + '_prep_t'], out[tag + '_prep_e'], out[tag + '_prep_b'] = p.tlm, p.elm, p.blm
    out[tag + '_dot'] = np.array([opfilt_tp.dot_op()(x(), r)])
    d = opfilt_tp.pre_op_diag(c['cls'], nf)(x())
    out[tag + '_prediag_t'], out[tag + '_prediag_e'], out[tag + '_prediag_b'] = d.tlm, d.elm, d.blm
    if tag == 'tp2':
        chain = multigrid.multigrid_chain(opfilt_tp, gi.chain_descr_tp(cd_solve), c['cls'], nf)
        sol = teblm([np.zeros(n, dtype=complex), np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)])
        trace = []
        ol = chain.log
        chain.log = lambda stage, it, eps, **kw_: (trace.append((stage.depth, it, eps)), ol(stage, it, eps, **kw_))
        chain.solve(sol, [c['tmap'], c['qmap'], c['umap']])
        out['tp2_soltn_t'], out['tp2_soltn_e'], out['tp2_soltn_b'] = sol.tlm, sol.elm, sol.blm
        out['tp2_trace'] = np.array([tr for tr in trace if tr[0] == 0])
        print('TP iterations (top level):', int(out['tp2_trace'][-1][1]))

fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_tp.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
