"""BASELINE.json configs[4]: high-resolution 'p' estimator (nside 4096, lmax_ivf 4000, lmax_qlm 5000) with every
transform of ONE estimate m-partitioned over the GPUs of the box.

    python scripts/bench_dist.py                     # 1 GPU: the plain single-GPU plan (reference point)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_dist.py [--check]             # N GPUs: dist_sht.DistPlan (NVLink peer stores + NCCL barriers)

Prints one JSON line (rank 0): seconds per estimate (CUDA events, max over ranks), algorithmic TFLOP/s summed over
GPUs (F0(4000) + 4 Fs(4000) + 2 Fs(5000) -- the reference runs two analyses, we merge them
into one, the credit stays at the reference's count only for the work we actually do: 1 analysis), and with
--check the relative L2 difference of the distributed qlm against the single-GPU plan on the same inputs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from plancklens_b200 import dist_sht, qest, sht  # noqa: E402
from plancklens_b200.helpers import mpi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nside', type=int, default=4096)
ap.add_argument('--lmax-ivf', type=int, default=4000)
ap.add_argument('--lmax-qlm', type=int, default=5000)
ap.add_argument('--steps', type=int, default=4)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--check', action='store_true')
a = ap.parse_args()

rank, size = mpi.init('nccl') if int(os.environ.get('WORLD_SIZE', 1)) > 1 else (0, 1)
if size == 1:
    torch.cuda.set_device(0)
nside, lmax, lmax_qlm = a.nside, a.lmax_ivf, a.lmax_qlm

# SO-like noise (n0s.get_N0 defaults: 1.4' beam, 5 / 5 sqrt(2) uK-arcmin); the same synthetic sky on every rank
bench.NLEV_T, bench.NLEV_P, bench.BEAM_AMIN = 5., 5. * np.sqrt(2.), 1.4
cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
tbar, ebar, bbar = [sht.dev_alm(x) for x in bench.filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl))]
cl_d = {k: sht.dev_fl(cls[k], lmax) for k in ('tt', 'ee', 'bb', 'te')}
twf = qest._combine(lmax, [(tbar, cl_d['tt']), (ebar, cl_d['te'])])
ewf = qest._combine(lmax, [(ebar, cl_d['ee']), (tbar, cl_d['te'])])
bwf = sht.almxfl(bbar, cl_d['bb'])

if size > 1:
    qe = qest.qe_device(nside, lmax, lmax_qlm, plan_ivf=dist_sht.DistPlan(nside, lmax),
                        plan_qlm=dist_sht.DistPlan(nside, lmax_qlm))
else:
    qe = qest.qe_device(nside, lmax, lmax_qlm)


def sync():
    if size > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


for _ in range(a.warmup):
    G, C = qe.p(tbar, ebar, bbar, twf, ewf, bwf)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = sht._lib.launch_count()
e0.record()
for _ in range(a.steps):
    G, C = qe.p(tbar, ebar, bbar, twf, ewf, bwf)
e1.record()
sync()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device='cuda')
if size > 1:
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
ms = float(t.item())
launches = (sht._lib.launch_count() - n0) // a.steps

stages = None
if size > 1:
    qe.plan_ivf.enable_timing(); qe.plan_qlm.enable_timing()
    qe.p(tbar, ebar, bbar, twf, ewf, bwf)
    stages = {'rank': rank, 'ivf': {k: round(v, 2) for k, v in qe.plan_ivf.stage_times().items()},
              'qlm': {k: round(v, 2) for k, v in qe.plan_qlm.stage_times().items()}}
    qe.plan_ivf.enable_timing(False); qe.plan_qlm.enable_timing(False)
    allst = [None] * size
    torch.distributed.all_gather_object(allst, stages)
    stages = allst

err = None
if a.check:
    ref = qest.qe_device(nside, lmax, lmax_qlm).p(tbar, ebar, bbar, twf, ewf, bwf)
    err = max(float(torch.linalg.norm(G - ref[0]) / torch.linalg.norm(ref[0])),
              float(torch.linalg.norm(C - ref[1]) / torch.linalg.norm(ref[1])))
    e = torch.tensor([err], dtype=torch.float64, device='cuda')
    if size > 1:
        torch.distributed.all_reduce(e, op=torch.distributed.ReduceOp.MAX)
    err = float(e.item())

if rank == 0:
    F0 = 8.0 * bench.n_lm(lmax) * 2 * nside
    Fs = 24.0 * bench.n_lm(lmax) * 2 * nside
    Fq = 24.0 * bench.n_lm(lmax_qlm) * 2 * nside
    flop = F0 + 4 * Fs + 1 * Fq
    print(json.dumps({"metric": "seconds per 'p' QE, one estimate m-partitioned over N GPUs", "n_gpus": size,
                      "nside": nside, "lmax_ivf": lmax, "lmax_qlm": lmax_qlm, "ms_per_qe": ms, "qlm_per_s": 1e3 / ms,
                      "algorithmic_tflops_total": flop / (ms * 1e-3) / 1e12, "launches_per_qe_per_rank": int(launches),
                      "rel_l2_vs_single_gpu": err, "stage_ms_one_estimate": stages, "steps": a.steps, "warmup": a.warmup,
                      "exchange": "peer stores over NVLink fused into legendre_synth / ring_anal kernels; "
                                  "2 NCCL one-element barriers per transform; qlm rows summed with one all-reduce"}))
if size > 1:
    mpi.finalize()
