"""Generates tests/golden/reference_golden_cinv.npz: the UNMODIFIED reference's config-3 filter libraries
(/root/reference/plancklens/filt/filt_cinv.py: cinv_t :74-203, cinv_p :223-338, library_cinv_sepTP :515-581) with their
DEFAULT multigrid chains (:113-116, :237-239) at the smallest size the constructors accept (nside 512, lmax 1024), on a
masked anisotropic-noise sky with monopole + dipole marginalisation -- run in the build container with
oracle/healpy_shim standing in for healpy (healpy is not installable here).

Pinned: top-level iteration counts, eps traces, the inverse-variance filtered alms out of `library_cinv_sepTP`
(i.e. `apply_ivf` after `rescal_cl`), the Wiener-filtered alms, the isotropic approximations (ftl / fel / fbl / tal) and
the mask the libraries derive.  Takes ~10 min on 8 cores (the dense coarse preconditioners are 4225 + 2178 operator
applications on the CPU oracle).

Run from the repo root:
  python tests/golden/make_golden_cinv.py            35 / 55 uK-arcmin  -> reference_golden_cinv.npz (T 12, P 3 iterations)
  python tests/golden/make_golden_cinv.py deep       3 / 4 uK-arcmin    -> reference_golden_cinv_deep.npz (T 8, P 22)
  python tests/golden/make_golden_cinv.py refresh    deep polarization filter pushed to eps_min = 1e-7 (the default chain
                                                     rows otherwise), past the `roundoff = 25` residual refresh of
                                                     cd_solve.py:79-81 -> reference_golden_cinv_refresh.npz
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'healpy_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import healpy as hp  # the shim  # noqa: E402
from plancklens.filt import filt_cinv  # noqa: E402  (reference)

import golden_inputs as gi  # noqa: E402

CLPATH = '/root/reference/plancklens/data/cls/FFP10_wdipole_lensedCls.dat'
MODE = sys.argv[1] if len(sys.argv) > 1 else ''
DEEP = MODE in ('deep', 'refresh')
c = gi.cinv_case(hp.alm2map, hp.alm2map_spin, CLPATH, **({'nlev_t': 3., 'nlev_p': 4.} if DEEP else {}))
lmax, nside = c['lmax'], c['nside']
out = {'mask_sum': np.array([c['mask'].sum()]), 'tmap_sum': np.array([c['tmap'].sum(), np.abs(c['tmap']).sum()]),
       'qmap_sum': np.array([c['qmap'].sum(), np.abs(c['qmap']).sum()])}


def traced(chain, store):
    orig = chain.log

    def log(stage, it, eps, **kw):
        store.append((stage.depth, it, eps))
        return orig(stage, it, eps, **kw)
    chain.log = log


if MODE == 'refresh':
    from plancklens.qcinv import cd_solve
    with tempfile.TemporaryDirectory() as tmp:
        descr = [[2, ["split(dense(), 32, diag_cl)"], 512, 256, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [1, ["split(stage(2),  512, diag_cl)"], 1024, 512, 3, 0.0, cd_solve.tr_cg, cd_solve.cache_mem()],
                 [0, ["split(stage(1), 1024, diag_cl)"], lmax, nside, np.inf, 1.0e-7, cd_solve.tr_cg, cd_solve.cache_mem()]]
        cinv_p = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, c['cls'], c['transf'], c['ninv_p'], chain_descr=descr)
        tr_p = []
        traced(cinv_p.chain, tr_p)
        elm, blm = cinv_p.apply_ivf([c['qmap'], c['umap']])
        out['p_trace'] = np.array([t for t in tr_p if t[0] == 0])
        print('P iterations to 1e-7:', int(out['p_trace'][-1][1]))
        for name, alm in (('elm', elm), ('blm', blm)):
            out[name + '_sample'] = gi.alm_sample(alm, lmax)
    fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_cinv_refresh.npz')
    np.savez_compressed(fn, **out)
    print('wrote', fn)
    sys.exit(0)

with tempfile.TemporaryDirectory() as tmp:
    t0 = time.time()
    cinv_t = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, c['cls'], c['transf'], c['ninv_t'],
                              marge_monopole=True, marge_dipole=True, marge_maps=[])
    cinv_p = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, c['cls'], c['transf'], c['ninv_p'])
    ivfs = filt_cinv.library_cinv_sepTP(os.path.join(tmp, 'ivfs'), gi.fixed_sim_lib(c), cinv_t, cinv_p, c['cls'])
    print('constructors: %.0f s' % (time.time() - t0))
    out['ftl'], out['fel'], out['fbl'] = ivfs.get_ftl(), ivfs.get_fel(), ivfs.get_fbl()
    out['tal_t'], out['tal_e'] = ivfs.get_tal('t'), ivfs.get_tal('e')
    out['fmask_sum'] = np.array([ivfs.get_fmask().sum()])

    tr_t, tr_p = [], []
    t0 = time.time()
    traced(cinv_t.chain, tr_t)       # util.jit: the chain (and its dense preconditioner) is built on first touch
    tlm = ivfs.get_sim_tlm(0)
    print('T solve (incl. dense preconditioner setup): %.0f s' % (time.time() - t0))
    t0 = time.time()
    traced(cinv_p.chain, tr_p)
    elm = ivfs.get_sim_elm(0)
    blm = ivfs.get_sim_blm(0)
    print('P solve (incl. dense preconditioner setup): %.0f s' % (time.time() - t0))
    out['t_trace'] = np.array([t for t in tr_t if t[0] == 0])
    out['p_trace'] = np.array([t for t in tr_p if t[0] == 0])
    out['t_trace_all'] = np.array(tr_t)
    out['p_trace_all'] = np.array(tr_p)
    print('T iterations:', int(out['t_trace'][-1][1]), ' P iterations:', int(out['p_trace'][-1][1]))
    for name, alm in (('tlm', tlm), ('elm', elm), ('blm', blm)):
        out[name + '_sample'] = gi.alm_sample(alm, lmax)
        out[name + '_cl'] = hp.alm2cl(alm)
        out[name + '_norm'] = np.array([np.linalg.norm(alm)])
    out['tmliklm_sample'] = gi.alm_sample(ivfs.get_sim_tmliklm(0), lmax)
    out['emliklm_sample'] = gi.alm_sample(ivfs.get_sim_emliklm(0), lmax)

    # second temperature solve: warm start from the first solution on a rescaled map (the `soltn` argument,
    # filt_cinv.py:196-203), through cinv_t.apply_ivf directly
    tr_t2 = []
    traced(cinv_t.chain, tr_t2)
    cl_resc = cinv_t.rescal_cl
    start = hp.almxfl(tlm, np.where(cl_resc > 0, 1. / np.where(cl_resc > 0, cl_resc, 1.), 0.))
    start = hp.almxfl(start, c['cls']['tt'] * cl_resc ** 2)    # back to the rescaled Wiener-filtered unknowns of the chain
    tlm2 = cinv_t.apply_ivf(1.05 * c['tmap'], soltn=start)
    out['t2_trace'] = np.array([t for t in tr_t2 if t[0] == 0])
    out['tlm2_sample'] = gi.alm_sample(tlm2, lmax)
    out['tlm2_cl'] = hp.alm2cl(tlm2)
    print('T warm-start iterations:', int(out['t2_trace'][-1][1]))

fn = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_cinv%s.npz' % ('_deep' if DEEP else ''))
np.savez_compressed(fn, **out)
print('wrote', fn, {k: np.shape(v) for k, v in out.items()})
