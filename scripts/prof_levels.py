"""Where the multigrid preconditioner of the masked-sky filters spends its time (nside 2048, lmax 2048, default chains):
every level of the stage tree is captured as its own CUDA graph and replayed alone, and the same number of trivial
dependent kernels is replayed as a graph to show the launch-gap floor.

  python scripts/prof_levels.py [--pol]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plancklens_b200 import hp, sht  # noqa: E402
from plancklens_b200.filt import filt_cinv  # noqa: E402
from plancklens_b200.qcinv import multigrid, util_alm  # noqa: E402

nside, lmax = 2048, int(os.environ.get('PLK_LMAX', 2048))
pol = '--pol' in sys.argv
cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
mask, z = bench.synthetic_sky_model(nside)
vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60
tmp = tempfile.mkdtemp(prefix='plk_levels_')
if pol:
    c = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, cls, transf, [[mask * (vamin / 55.) ** 2 * (1 + 0.5 * z ** 2)]])
else:
    c = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, cls, transf, [mask * (vamin / 35.) ** 2 * (1 + 0.5 * z ** 2)],
                         marge_monopole=True, marge_dipole=True)
rng = np.random.default_rng(0)


def vec(l):
    def one():
        a = rng.standard_normal(sht.alm_size(l)) + 1j * rng.standard_normal(sht.alm_size(l))
        return util_alm.dalm(sht.dev_alm(a))
    return util_alm.eblm([one(), one()]) if pol else one()


def timed(op, v, n=10):
    g = multigrid.graphed_op(op)
    n0 = sht._lib.launch_count()
    g(v)
    nl = sht._lib.launch_count() - n0
    g(v); g(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, nl


top = c.chain.bstage.pre_ops[0]
cur = getattr(top, 'op', top)
rows = []
while cur is not None:
    name = type(cur).__name__
    l = getattr(cur, 'lmax')
    ms, nl = timed(cur, vec(l))
    extra = ''
    if isinstance(cur, multigrid.pre_op_multigrid):
        f_ms, f_nl = timed(cur.fwd_op, vec(l))
        extra = '  (one fwd_op alone: %.3f ms, %d launches; nside %d, %d iterations)' % (f_ms, f_nl, cur.nside, cur.iter_max)
    rows.append((name, l, ms, nl))
    print('%-18s lmax %5d : %8.3f ms  %5d launches%s' % (name, l, ms, nl, extra), file=sys.stderr)
    if isinstance(cur, multigrid.pre_op_split):
        cur = cur.pre_op_low
    elif isinstance(cur, multigrid.pre_op_multigrid):
        cur = cur.pre_ops[0]
    else:
        cur = None

# launch-gap floor: a chain of dependent trivial kernels in one graph
x = torch.zeros(66, dtype=torch.complex128, device='cuda')
fl = torch.ones(11, dtype=torch.float64, device='cuda')
for nk in (200, 600):
    g = torch.cuda.CUDAGraph()
    sht.almxfl(x, fl, out=x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, capture_error_mode='thread_local'):
        for _ in range(nk):
            sht.almxfl(x, fl, out=x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print('graph of %d dependent one-block kernels: %.3f ms per replay = %.2f us per node' % (nk, e0.elapsed_time(e1) / 10, e0.elapsed_time(e1) / 10 / nk * 1e3), file=sys.stderr)
