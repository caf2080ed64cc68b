"""Generates tests/golden/reference_golden.npz by running the UNMODIFIED reference Python
(/root/reference/plancklens) in the build container, with oracle/healpy_shim standing in for healpy
(healpy itself is not installed).

What this pins: everything the reference computes above the SHT seam -- qcinv operators (opfilt_tt, opfilt_pp),
cd_solve iteration counts and residual traces, multigrid / dense preconditioners, the qest fast path for
'ptt', 'p_p', 'p' and the generic utils_qe.qe_eval path -- on small seeded inputs.
What it does not pin: the transforms themselves (the shim calls the oracle).

Run from the repo root:  python tests/golden/make_golden.py
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
from plancklens import qest, utils  # noqa: E402  (reference)
from plancklens.qcinv import cd_solve, multigrid, opfilt_pp, opfilt_tt  # noqa: E402  (reference)
from plancklens.qcinv.util_alm import eblm  # noqa: E402

import golden_inputs as gi  # noqa: E402

out = {}

# ------------------------------------------------------------------ CG, temperature
c = gi.cg_case()
ninv_filt = opfilt_tt.alm_filter_ninv(c['ninv_t'], c['transf'], marge_monopole=True, marge_dipole=True)
fwd = opfilt_tt.fwd_op(c['cls'], ninv_filt)
out['tt_fwd'] = fwd(c['x_t'].copy())
out['tt_prep'] = opfilt_tt.calc_prep(c['tmap'], c['cls'], ninv_filt)
out['tt_prediag'] = opfilt_tt.pre_op_diag(c['cls'], ninv_filt)(c['x_t'].copy())
out['tt_dot'] = np.array([opfilt_tt.dot_op()(c['x_t'], out['tt_fwd'])])
am = c['x_t'].copy()
ninv_filt.apply_alm(am)
out['tt_apply_alm'] = am
chain = multigrid.multigrid_chain(opfilt_tt, gi.chain_descr_t(cd_solve), c['cls'], ninv_filt)
sol = np.zeros(hp.Alm.getsize(c['lmax']), dtype=complex)
trace = []
orig_log = chain.log
chain.log = lambda stage, it, eps, **kw: (trace.append((stage.depth, it, eps)), orig_log(stage, it, eps, **kw))
chain.solve(sol, c['tmap'])
out['tt_soltn'] = sol
out['tt_trace'] = np.array([t for t in trace if t[0] == 0])
print('T iterations (top level):', int(out['tt_trace'][-1][1]))

# diagonal preconditioner only (what oracle/ref_cg.pcg restates)
chain_d = multigrid.multigrid_chain(opfilt_tt, [[0, ["diag_cl"], c['lmax'], c['nside'], np.inf, 1.0e-6, cd_solve.tr_cg, cd_solve.cache_mem()]], c['cls'], ninv_filt)
sold = np.zeros(hp.Alm.getsize(c['lmax']), dtype=complex)
traced = []
chain_d.log = lambda stage, it, eps, **kw: traced.append((stage.depth, it, eps))
chain_d.solve(sold, c['tmap'])
out['tt_diag_soltn'] = sold
out['tt_diag_trace'] = np.array(traced)
print('T iterations (diag only):', int(out['tt_diag_trace'][-1][1]))

# ------------------------------------------------------------------ CG, polarization
for tag, ninv_p in (('pp', c['ninv_p1']), ('pp3', c['ninv_p3'])):
    nf = opfilt_pp.alm_filter_ninv(ninv_p, c['transf'])
    fwdp = opfilt_pp.fwd_op(c['cls'], nf)
    x = eblm([c['x_e'].copy(), c['x_b'].copy()])
    r = fwdp(x)
    out[tag + '_fwd_e'], out[tag + '_fwd_b'] = r.elm, r.blm
    r = opfilt_pp.calc_prep([c['qmap'], c['umap']], c['cls'], nf)
    out[tag + '_prep_e'], out[tag + '_prep_b'] = r.elm, r.blm
    out[tag + '_dot'] = np.array([opfilt_pp.dot_op()(x, fwdp(x))])
    if tag == 'pp':
        r = opfilt_pp.pre_op_diag(c['cls'], nf)(x)
        out['pp_prediag_e'], out['pp_prediag_b'] = r.elm, r.blm
        chainp = multigrid.multigrid_chain(opfilt_pp, gi.chain_descr_p(cd_solve), c['cls'], nf)
        solp = eblm([np.zeros(hp.Alm.getsize(c['lmax']), dtype=complex), np.zeros(hp.Alm.getsize(c['lmax']), dtype=complex)])
        tracep = []
        ol = chainp.log
        chainp.log = lambda stage, it, eps, **kw: (tracep.append((stage.depth, it, eps)), ol(stage, it, eps, **kw))
        chainp.solve(solp, [c['qmap'], c['umap']])
        out['pp_soltn_e'], out['pp_soltn_b'] = solp.elm, solp.blm
        out['pp_trace'] = np.array([t for t in tracep if t[0] == 0])
        print('P iterations (top level):', int(out['pp_trace'][-1][1]))

# ------------------------------------------------------------------ QE fast path and generic path
q = gi.qe_case()


class mem_ivfs:
    """In-memory filtering library with the duck-type qest needs."""
    lib_dir = None

    def __init__(self, q, tag):
        self.q, self.tag = q, tag

    def hashdict(self):
        return {'tag': self.tag}

    def get_fmask(self):
        return np.ones(hp.nside2npix(self.q['nside']))

    def get_sim_tlm(self, idx): return self.q['tlm' + self.tag].copy()
    def get_sim_elm(self, idx): return self.q['elm' + self.tag].copy()
    def get_sim_blm(self, idx): return self.q['blm' + self.tag].copy()
    def get_sim_tmliklm(self, idx): return hp.almxfl(self.get_sim_tlm(idx), self.q['cls']['tt'])
    def get_sim_emliklm(self, idx): return hp.almxfl(self.get_sim_elm(idx), self.q['cls']['ee'])
    def get_sim_bmliklm(self, idx): return hp.almxfl(self.get_sim_blm(idx), self.q['cls']['bb'])


with tempfile.TemporaryDirectory() as tmp:
    iv1, iv2 = mem_ivfs(q, '1'), mem_ivfs(q, '2')
    lib_dd = qest.library_sepTP(os.path.join(tmp, 'dd'), iv1, iv1, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
    lib_ds = qest.library_sepTP(os.path.join(tmp, 'ds'), iv1, iv2, q['cls']['te'], q['nside'], lmax_qlm=q['lmax_qlm'])
    for k in ['ptt', 'p_p', 'p']:
        out['qe_dd_' + k] = lib_dd.get_sim_qlm(k, 0)
        out['qe_dd_x' + k[1:]] = lib_dd.get_sim_qlm('x' + k[1:], 0)
        out['qe_ds_' + k] = lib_ds.get_sim_qlm(k, 0)
    # generic path (qest.eval_qe -> qresp.get_qes -> utils_qe.qe_eval)
    get_alm = lambda a: {'t': q['tlm1'], 'e': q['elm1'], 'b': q['blm1']}[a].copy()
    for k in ['ptt', 'p_p']:
        G, C = qest.eval_qe(k, q['lmax'], q['cls'], get_alm, q['nside'], q['lmax_qlm'], verbose=False)
        out['qe_gen_' + k] = G
        out['qe_gen_x' + k[1:]] = C if np.ndim(C) else np.zeros_like(G)

fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz')
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
