"""Masked-sky CG inverse-variance filter at full size (BASELINE.json configs[2]): cinv_t / cinv_p with the
reference's default multigrid chains (filt_cinv.py:113-116, :237-239), synthetic Galactic mask + 2000 point-source
discs + anisotropic noise, eps_min = 1e-5.

`run()` returns, per field, the top-level iteration count, seconds and iterations per second of
  * `e2e`: `cinv_t.apply_ivf(numpy map) -> numpy alm` (pageable host arrays in and out, second solve of the chain),
  * `device_resident`: `chain.solve` on a map already in HBM,
plus the CUDA-event time of one top-level forward operator and one (CUDA-graph replayed) preconditioner.
bench.py embeds this in its JSON line; run as a script for the numbers alone.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def synthetic_mask(nside, rng):
    """|b| < 20 deg cut + 2000 discs of radius 10 arcmin, from ring geometry on the host."""
    from golden_inputs import pix_z
    npix = 12 * nside ** 2
    z = pix_z(nside)
    mask = (np.abs(z) >= np.sin(np.deg2rad(20.))).astype(float)
    zc = rng.uniform(-1, 1, 2000)
    pc = rng.uniform(0, 2 * np.pi, 2000)
    phi = np.empty(npix)
    p = 0
    for i in range(1, 4 * nside):
        ii = i if i < nside else (nside if i <= 3 * nside else 4 * nside - i)
        n = 4 * ii
        ph0 = np.pi / n if (i < nside or i > 3 * nside or (i - nside) % 2 == 0) else 0.0
        phi[p:p + n] = ph0 + 2 * np.pi * np.arange(n) / n
        p += n
    s = np.sqrt(1 - z * z)
    vec = np.stack([s * np.cos(phi), s * np.sin(phi), z], 1)
    cr = np.cos(np.deg2rad(10. / 60.))
    order = np.argsort(z)
    zs = z[order]
    for k in range(2000):
        sc = np.sqrt(1 - zc[k] ** 2)
        c = np.array([sc * np.cos(pc[k]), sc * np.sin(pc[k]), zc[k]])
        lo, hi = np.searchsorted(zs, [zc[k] - 0.004, zc[k] + 0.004])
        cand = order[lo:hi]
        mask[cand[vec[cand] @ c > cr]] = 0.0
    return mask, z


def _op_times(chain, v, profile=False):
    """device ms of one top-level forward operator and one preconditioner application"""
    fwd = chain.opfilt.fwd_op(chain.s_cls, chain.n_inv_filt)
    if profile:     # ncu --profile-from-start off: one eager application of the top-level preconditioner
        op = chain.bstage.pre_ops[0]
        op = getattr(op, 'op', op)
        op(v); torch.cuda.synchronize()
        torch.cuda.profiler.start()
        op(v); torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return {}
    out = {}
    for name, op in (('fwd_op_ms', fwd), ('pre_op_ms', chain.bstage.pre_ops[0])):
        op(v); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            op(v)
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 3
    return out


def run(nside=2048, lmax=2048, do_t=True, do_p=True, profile_pre_op=False, verbose=True):
    import bench
    from plancklens_b200 import hp, sht
    from plancklens_b200.filt import filt_cinv
    from plancklens_b200.qcinv import util_alm
    cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
    mask, z = synthetic_mask(nside, np.random.default_rng(7))
    vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60
    ninv_t = mask * (vamin / bench.NLEV_T) ** 2 * (1 + 0.5 * z ** 2)
    ninv_p = mask * (vamin / bench.NLEV_P) ** 2 * (1 + 0.5 * z ** 2)
    alms = bench.filtered_sim(0, lmax, cls, transf, (np.ones(lmax + 1),) * 3)   # unfiltered a + n / b
    out = {'nside': nside, 'lmax': lmax, 'fsky': float(mask.mean()), 'eps_min': 1e-5}
    tmp = tempfile.mkdtemp(prefix='plk_cg_')

    def solve_twice(make, data, dev_data, zeros):
        t0 = time.time()
        cinv = make()
        _ = cinv.chain.bstage          # instantiate: degraded filters, dense preconditioner, stages
        torch.cuda.synchronize()
        t_setup = time.time() - t0
        cinv.apply_ivf(data)           # first solve: plans, tables, CUDA-graph capture of the preconditioner
        torch.cuda.synchronize()
        n0 = sht._lib.launch_count(); t0 = time.time()
        res = cinv.apply_ivf(data)
        torch.cuda.synchronize(); dt = time.time() - t0
        launches = sht._lib.launch_count() - n0
        it = cinv.chain.niter
        sol = zeros()
        torch.cuda.synchronize(); t1 = time.time()
        cinv.chain.solve(sol, dev_data)
        torch.cuda.synchronize(); dt_dev = time.time() - t1
        r = {'iterations': int(it), 'final_eps': float(cinv.chain.last_monitor.trace[-1][1]), 'setup_s': t_setup,
             'e2e': {'seconds': dt, 'iter_per_s': it / dt}, 'device_resident': {'seconds': dt_dev, 'iter_per_s': it / dt_dev},
             'eager_launches_per_solve': int(launches)}
        return r, res, cinv

    if do_t:
        tmap = hp.alm2map(hp.almxfl(alms[0], transf), nside)
        mk = lambda: filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, cls, transf, [ninv_t], marge_monopole=True, marge_dipole=True)
        r, tlm, cinv_t = solve_twice(mk, tmap, sht.dev_map(tmap), lambda: util_alm.dalm.zeros(lmax))
        r.update(_op_times(cinv_t.chain, util_alm.dalm(sht.dev_alm(tlm)), profile_pre_op))
        out['T'] = r
        if verbose:
            print('CG-T', r)
    if do_p:
        qmap, umap = hp.alm2map_spin([hp.almxfl(alms[1], transf), hp.almxfl(alms[2], transf)], nside, 2, lmax)
        mk = lambda: filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, cls, transf, [[ninv_p]])
        r, (elm, blm), cinv_p = solve_twice(mk, [qmap, umap], [sht.dev_map(qmap), sht.dev_map(umap)],
                                            lambda: util_alm.eblm([util_alm.dalm.zeros(lmax), util_alm.dalm.zeros(lmax)]))
        r.update(_op_times(cinv_p.chain, util_alm.eblm([util_alm.dalm(sht.dev_alm(elm)), util_alm.dalm(sht.dev_alm(blm))])))
        out['P'] = r
        if verbose:
            print('CG-P', r)
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--nside', type=int, default=2048)
    ap.add_argument('--lmax', type=int, default=2048)
    ap.add_argument('--pol', action='store_true')
    ap.add_argument('--skip-t', action='store_true')
    ap.add_argument('--profile-pre-op', action='store_true')
    a = ap.parse_args()
    print(json.dumps({'cg': run(a.nside, a.lmax, not a.skip_t, a.pol, a.profile_pre_op)}))
