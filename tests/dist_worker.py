"""torchrun worker of tests/test_dist.py::test_real_processes_two_gpus: DistPlan against the single-GPU plan."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from helpers import rand_alm  # noqa: E402
from plancklens_b200 import dist_sht, sht  # noqa: E402
from plancklens_b200.helpers import mpi  # noqa: E402

rank, size = mpi.init('nccl')
for nside, lmax in ((64, 150), (512, 700)):
    rng = np.random.default_rng(5)     # same inputs on every rank
    plan = sht.get_plan(nside, lmax)
    dp = dist_sht.DistPlan(nside, lmax)
    ranges = dp.pixel_ranges()

    def own(t):
        return torch.cat([t[lo:hi] for lo, hi in ranges])

    a = sht.dev_alm(rand_alm(rng, lmax))
    g, c = sht.dev_alm(rand_alm(rng, lmax, 2)), sht.dev_alm(rand_alm(rng, lmax, 2))
    m1 = sht.dev_map(rng.standard_normal(12 * nside ** 2)); m2 = sht.dev_map(rng.standard_normal(12 * nside ** 2))
    for rep in range(3):     # repeated calls exercise the write-after-read barriers
        assert torch.equal(own(dp.alm2map(a)), own(plan.alm2map(a)))
        got, ref = dp.alm2map_spin(g, c, 2), plan.alm2map_spin(g, c, 2)
        assert torch.equal(own(got[0]), own(ref[0])) and torch.equal(own(got[1]), own(ref[1]))
        # analysis: the all-reduce adds exact zeros to each rank's rows -> bit-identical
        assert torch.equal(dp.map2alm(m1), plan.map2alm(m1))
        got, ref = dp.map2alm_spin(m1, m2, 1), plan.map2alm_spin(m1, m2, 1)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
torch.cuda.synchronize()
torch.distributed.barrier()
if rank == 0:
    print('DIST OK on %d ranks' % size)
mpi.finalize()
