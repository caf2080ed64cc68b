"""Launch list of ONE application of the top-level multigrid preconditioner of the masked-sky filters (nside 2048,
lmax 2048, default chains), for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python scripts/prof_preop.py [--pol]

The preconditioner is applied eagerly between cudaProfilerStart / Stop (kernel nodes of a replayed CUDA graph are the
same launches); the CUDA-event time of the graph replay the solver uses is printed beside it."""
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
from plancklens_b200.qcinv import util_alm  # noqa: E402

nside, lmax = 2048, int(os.environ.get('PLK_LMAX', 2048))
pol = '--pol' in sys.argv
cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
mask, z = bench.synthetic_sky_model(nside)
vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60
tmp = tempfile.mkdtemp(prefix='plk_preop_')
rng = np.random.default_rng(1)
if pol:
    c = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, cls, transf, [[mask * (vamin / 55.) ** 2 * (1 + 0.5 * z ** 2)]])
    v = util_alm.eblm([util_alm.dalm(sht.dev_alm(x)) for x in bench.filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl))[1:]])
else:
    c = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, cls, transf, [mask * (vamin / 35.) ** 2 * (1 + 0.5 * z ** 2)],
                         marge_monopole=True, marge_dipole=True)
    v = util_alm.dalm(sht.dev_alm(bench.filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl))[0]))
op = c.chain.bstage.pre_ops[0]
eager = getattr(op, 'op', op)
for _ in range(3):
    op(v)                      # eager, capture, replay
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    op(v)
e1.record()
torch.cuda.synchronize()
fwd = c.chain.opfilt.fwd_op(c.chain.s_cls, c.chain.n_inv_filt)
fwd(v); torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(5):
    fwd(v)
f1.record()
torch.cuda.synchronize()
n0 = sht._lib.launch_count()
torch.cuda.profiler.start()
eager(v)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('pre_op graph replay %.3f ms, fwd_op %.3f ms, launches in one eager pre_op: %d' %
      (e0.elapsed_time(e1) / 5, f0.elapsed_time(f1) / 5, sht._lib.launch_count() - n0), file=sys.stderr)
