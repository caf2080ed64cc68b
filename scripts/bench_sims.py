"""BASELINE.json configs[3]: mean-field / N0 style batch of masked-sky 'p' estimates, simulations sharded over the GPUs
of one box (idx % N == rank, the reference's `jobs[mpi.rank::mpi.size]` pattern of examples/run_qlms.py:72).

Every simulation goes through the reference-shaped pipeline of params/anisofilt_example.py:
  sims.get_sim_tmap / get_sim_pmap (Gaussian CMB + white noise, synthesis on the GPU)
  -> cinv_t / cinv_p multigrid-preconditioned CG (eps 1e-5, default chains)
  -> qlms_dd.get_sim_qlm('p', idx) (cached on disk under the reference's file names).
The host-side random draws of simulation i + 1 (numpy, 3 x 50 M normals) are prefetched on worker threads while the
GPU filters simulation i; everything else is the library's own code path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_sims.py --per-rank 6
Prints one JSON line (rank 0): aggregate simulations per second (wall clock, max over ranks, after one warm-up
simulation per rank that also builds the dense preconditioners and captures the CUDA graphs).
"""
import argparse
import concurrent.futures as cf
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument('--nside', type=int, default=2048)
ap.add_argument('--lmax', type=int, default=2048)
ap.add_argument('--per-rank', type=int, default=6, help="timed simulations per rank (weak scaling)")
a = ap.parse_args()

from plancklens_b200.helpers import mpi  # noqa: E402
rank, size = mpi.init('nccl') if int(os.environ.get('WORLD_SIZE', 1)) > 1 else (0, 1)
if size == 1:
    torch.cuda.set_device(0)

os.environ.setdefault('PLK_CACHE_FORMAT', 'npy')     # .npy caches under the reference's file names (FITS: +0.15 s per alm)
os.environ['PLENS'] = os.path.join('/tmp', 'plk_cfg4_%s_%d' % (os.environ.get('MASTER_PORT', 'single'), os.getppid() if size > 1 else os.getpid()))
os.environ.update({'PLK_NSIDE': str(a.nside), 'PLK_LMAX_IVF': str(a.lmax), 'PLK_LMAX_QLM': str(a.lmax), 'PLK_NSIMS': '320'})
spec = importlib.util.spec_from_file_location('anisofilt_example', os.path.join(ROOT, 'params', 'anisofilt_example.py'))
par = importlib.util.module_from_spec(spec)
sys.stdout = sys.stderr          # the libraries log to stdout; keep it for the JSON line
spec.loader.exec_module(par)


class prefetched:
    """get_sim(idx, idf) of a phase library with the draws computed ahead of time on worker threads"""

    def __init__(self, lib, pool):
        self.lib, self.pool, self.fut = lib, pool, {}
        self.nfields = lib.nfields
        for name in ('shape', 'lmax'):
            if hasattr(lib, name):
                setattr(self, name, getattr(lib, name))

    def request(self, idx):
        for idf in range(self.lib.nfields):
            if (idx, idf) not in self.fut:
                self.fut[(idx, idf)] = self.pool.submit(self.lib.get_sim, idx, idf)

    def get_sim(self, idx, idf=None, phas_only=False):
        assert idf is not None
        f = self.fut.pop((idx, idf), None)
        return f.result() if f is not None else self.lib.get_sim(idx, idf)

    def hashdict(self):
        return self.lib.hashdict()

    def is_full(self):
        return True


pool = cf.ThreadPoolExecutor(max_workers=3)
maps_lib = par.sims.sim_lib if hasattr(par.sims, 'sim_lib') else par.sims
pix = prefetched(maps_lib.pix_lib_phas, pool)
maps_lib.pix_lib_phas = pix
maps_lib.device_maps = True       # simulated maps stay on the GPU on their way into the CG filters
cmb = prefetched(par.cmb_sims.lib_pha, pool) if hasattr(par.cmb_sims, 'lib_pha') else None
if cmb is not None:
    par.cmb_sims.lib_pha = cmb

idxs = list(range(rank, size * (a.per_rank + 1), size))     # first one is the warm-up


def request(idx):
    pix.request(idx)
    if cmb is not None:
        cmb.request(idx)


times = []
request(idxs[0])
for i, idx in enumerate(idxs):
    if i + 1 < len(idxs):
        request(idxs[i + 1])
    if i == 1:
        if size > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t_start = time.perf_counter()
    t0 = time.perf_counter()
    G = par.qlms_dd.get_sim_qlm('p', idx)
    times.append(time.perf_counter() - t0)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t_start], dtype=torch.float64, device='cuda')
if size > 1:
    torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
assert np.all(np.isfinite(G)) and np.any(G != 0)
sys.stdout = sys.__stdout__
if rank == 0:
    print(json.dumps({"metric": "masked-sky 'p' QE simulations per second (CG-filtered T and P, sims sharded over GPUs)",
                      "n_gpus": size, "sims_timed": size * a.per_rank, "seconds": float(dt.item()),
                      "sims_per_s": size * a.per_rank / float(dt.item()), "nside": a.nside, "lmax": a.lmax,
                      "cg_iterations": {"T": int(par.cinv_t.chain.niter), "P": int(par.cinv_p.chain.niter)},
                      "rank0_seconds_per_sim": [round(t, 3) for t in times[1:]], "rank0_warmup_s": round(times[0], 2),
                      "scaling": "weak"}))
if size > 1:
    mpi.finalize()
