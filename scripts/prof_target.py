"""Host-side breakdown of one simulation of the north_star target pipeline (bench.py run_target): wall time of each
library call with a device synchronisation after it, next to the CUDA-event time of the whole simulation without the
synchronisations.  python scripts/prof_target.py [lmax]"""
import contextlib
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plancklens_b200 import sht  # noqa: E402

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
os.environ.setdefault('PLK_CACHE_FORMAT', 'npy')
tmp = tempfile.mkdtemp(prefix='plk_proft_')
with contextlib.redirect_stdout(open(os.devnull, 'w')):
    mask, z = bench.synthetic_sky_model(bench.NSIDE)
    lib = bench.build_target(lmax, tmp, mask, z)
    q = lib['qlms_dd']
    for i in range(2):
        q.get_sim_qlm_dev('p', i)
    torch.cuda.synchronize()

    def timed(label, fn, acc):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        acc.append((label, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
        return r
    rows = []
    for idx in (10, 11):
        acc = []
        raw = lib['ivfs_raw']
        tmap = timed('get_sim_tmap_dev', lambda: lib['sims'].get_sim_tmap_dev(idx), acc)
        pm = timed('get_sim_pmap_dev', lambda: lib['sims'].get_sim_pmap_dev(idx), acc)
        tlm = timed('cinv_t.apply_ivf_dev', lambda: lib['cinv_t'].apply_ivf_dev(tmap), acc)
        eb = timed('cinv_p.apply_ivf_dev', lambda: lib['cinv_p'].apply_ivf_dev(list(pm)), acc)
        timed('raw.get_sim_teblm_dev (full path, new idx)', lambda: raw.get_sim_teblm_dev(idx + 100), acc)
        timed('qlms_dd.get_sim_qlm_dev (ivfs cached)', lambda: q.get_sim_qlm_dev('p', idx + 100), acc)
        rows.append(acc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for idx in (20, 21, 22):
        q.get_sim_qlm_dev('p', idx)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
for acc in rows:
    for label, host, total in acc:
        print('%-48s host %8.1f ms   host+device %8.1f ms' % (label, host, total))
    print()
print('3 simulations back to back: %.1f ms per simulation (events), %.1f (wall)' % (e0.elapsed_time(e1) / 3, 1e3 * wall / 3))
print('CG iterations T %d P %d' % (lib['cinv_t'].chain.niter, lib['cinv_p'].chain.niter))
