"""Where the set-up time of the masked-sky filters goes (cinv_t / cinv_p chains at nside 2048): cProfile of the
constructors + first touch of the chain, top cumulative entries.  python scripts/prof_setup.py [--pol]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import bench_cg  # noqa: E402
from plancklens_b200 import hp  # noqa: E402
from plancklens_b200.filt import filt_cinv  # noqa: E402

nside, lmax = 2048, 2048
pol = '--pol' in sys.argv
cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
t0 = time.time()
mask, z = bench_cg.synthetic_mask(nside, np.random.default_rng(7))
print('mask: %.1f s' % (time.time() - t0))
vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60
tmp = tempfile.mkdtemp(prefix='plk_setup_')
pr = cProfile.Profile()
pr.enable()
t0 = time.time()
if pol:
    c = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, cls, transf, [[mask * (vamin / 55.) ** 2 * (1 + 0.5 * z ** 2)]])
else:
    c = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, cls, transf, [mask * (vamin / 35.) ** 2 * (1 + 0.5 * z ** 2)],
                         marge_monopole=True, marge_dipole=True)
_ = c.chain.bstage
torch.cuda.synchronize()
pr.disable()
print('setup: %.1f s' % (time.time() - t0))
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
