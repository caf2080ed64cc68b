"""torchrun worker of tests/test_dist.py::test_m_distributed_cg_two_gpus: the masked-sky T and P filters with the
forward operator m-partitioned (qcinv/dist_cg.py) against the single-GPU solve of the same chain."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import golden_inputs as gi  # noqa: E402
from oracle import ref_sht  # noqa: E402
from plancklens_b200.filt import filt_cinv  # noqa: E402
from plancklens_b200.helpers import mpi  # noqa: E402
from plancklens_b200.qcinv import dist_cg, util_alm  # noqa: E402

rank, size = mpi.init('nccl')
CLPATH = os.path.join(ROOT, 'plancklens_b200', 'data', 'cls', 'FFP10_wdipole_lensedCls.dat')
c = gi.cinv_case(ref_sht.alm2map, ref_sht.alm2map_spin, CLPATH)       # same inputs on every rank
tmp = mpi.bcast(tempfile.mkdtemp(prefix='plk_distcg_') if rank == 0 else None)
sys.stdout = open(os.devnull, 'w') if rank else sys.stdout
lmax, nside = c['lmax'], c['nside']
cinv_t = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, c['cls'], c['transf'], c['ninv_t'],
                          marge_monopole=True, marge_dipole=True, marge_maps=[])
cinv_p = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, c['cls'], c['transf'], c['ninv_p'])

for name, chain, maps, zero in (
        ('T', cinv_t.chain, c['tmap'], lambda: util_alm.dalm.zeros(lmax)),
        ('P', cinv_p.chain, [c['qmap'], c['umap']], lambda: util_alm.eblm([util_alm.dalm.zeros(lmax), util_alm.dalm.zeros(lmax)]))):
    ref = zero()
    chain.solve(ref, maps)                       # every rank on its own GPU: the single-GPU reference
    n_ref, tr_ref = chain.niter, np.array([t[1] for t in chain.last_monitor.trace])
    dc = dist_cg.dist_chain(chain)
    for rep in range(2):
        got = zero()
        n = dc.solve(got, maps)
        tr = np.array([t[1] for t in chain.last_monitor.trace])
        assert n == n_ref, (name, n, n_ref)
        assert np.allclose(tr, tr_ref, rtol=1e-6, atol=0), (name, tr, tr_ref)
        for a, b in zip(dist_cg._comps(got), dist_cg._comps(ref)):
            err = float(torch.linalg.norm(a.t - b.t) / torch.linalg.norm(b.t))
            assert err < 1e-8, (name, err)
    # every rank holds the same replicated solution
    chk = torch.stack([torch.linalg.norm(x.t) for x in dist_cg._comps(got)])
    allc = [torch.empty_like(chk) for _ in range(size)]
    torch.distributed.all_gather(allc, chk)
    assert all(torch.equal(allc[0], x) for x in allc), name
    if rank == 0:
        sys.__stdout__.write('%s: %d iterations on %d ranks, same trace as one GPU\n' % (name, n, size))
torch.cuda.synchronize()
torch.distributed.barrier()
if rank == 0:
    sys.__stdout__.write('DIST CG OK on %d ranks\n' % size)
mpi.finalize()
