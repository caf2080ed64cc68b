"""Generates tests/golden/reference_golden_kk.npz with the UNMODIFIED reference (see make_golden.py): the
convergence-map filter qcinv/opfilt_kk.py -- operators and a two-level multigrid solve.
Run from the repo root:  python tests/golden/make_golden_kk.py
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
from plancklens.qcinv import cd_solve, multigrid, opfilt_kk  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

c = gi.cg_case()
s_cls = gi.kk_cls(c['lmax'])
out = {}
nf = opfilt_kk.alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
fwd = opfilt_kk.fwd_op(s_cls, nf)
out['kk_fwd'] = fwd(c['x_t'].copy())
out['kk_prep'] = opfilt_kk.calc_prep(c['tmap'], s_cls, nf)
out['kk_prediag'] = opfilt_kk.pre_op_diag(s_cls, nf)(c['x_t'].copy())
out['kk_dot'] = np.array([opfilt_kk.dot_op()(c['x_t'], out['kk_fwd'])])
chain = multigrid.multigrid_chain(opfilt_kk, gi.chain_descr_t(cd_solve), s_cls, nf)
sol = np.zeros(hp.Alm.getsize(c['lmax']), dtype=complex)
trace = []
ol = chain.log
chain.log = lambda stage, it, eps, **kw: (trace.append((stage.depth, it, eps)), ol(stage, it, eps, **kw))
chain.solve(sol, c['tmap'])
out['kk_soltn'] = sol
out['kk_trace'] = np.array([t for t in trace if t[0] == 0])
print('kk iterations (top level):', int(out['kk_trace'][-1][1]))
fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_kk.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
