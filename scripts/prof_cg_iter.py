"""CUDA-event breakdown of ONE top-level CG iteration of the masked-sky filters (cd_solve.cd_solve_dev) at nside 2048:
python scripts/prof_cg_iter.py [--pol] [lmax]"""
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
from plancklens_b200.qcinv import cd_solve, util_alm  # noqa: E402

pol = '--pol' in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith('--')]
nside, lmax = 2048, int(args[0]) if args else 2048
cls, transf, ftl, fel, fbl = bench.fiducial(lmax)
mask, z = bench.synthetic_sky_model(nside)
vamin = np.sqrt(hp.nside2pixarea(nside, degrees=True)) * 60
tmp = tempfile.mkdtemp(prefix='plk_cgit_')
if pol:
    c = filt_cinv.cinv_p(os.path.join(tmp, 'cinv_p'), lmax, nside, cls, transf, [[mask * (vamin / 55.) ** 2 * (1 + 0.5 * z ** 2)]])
    v = util_alm.eblm([util_alm.dalm(sht.dev_alm(x)) for x in bench.filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl))[1:]])
else:
    c = filt_cinv.cinv_t(os.path.join(tmp, 'cinv_t'), lmax, nside, cls, transf, [mask * (vamin / 35.) ** 2 * (1 + 0.5 * z ** 2)],
                         marge_monopole=True, marge_dipole=True)
    v = util_alm.dalm(sht.dev_alm(bench.filtered_sim(0, lmax, cls, transf, (ftl, fel, fbl))[0]))
chain = c.chain
with sht.use_lane(getattr(c, 'lane', 0)):
    pre_op = chain.bstage.pre_ops[0]
    fwd_op = chain.opfilt.fwd_op(chain.s_cls, chain.n_inv_filt)
    dot_op = chain.opfilt.dot_op()
    x = v * 0.0
    residual = v.copy()
    for _ in range(3):
        d = pre_op(residual)
    labels = ['fwd_op(d)', 'dot d.r', 'dot d.Ad + alpha', 'update x, r', 'pre_op(r)', 'dot dn.Ad + beta', 'd = dn + beta d', 'monitor: dot r.r + item()']
    tot = np.zeros(len(labels))
    nit = 6
    for it in range(nit):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(labels) + 1)]
        ev[0].record()
        Ad = fwd_op(d); ev[1].record()
        delta = dot_op.fused(d, residual); ev[2].record()
        t3 = dot_op.fused(d, Ad, num=delta[0:1]); ev[3].record()
        cd_solve._update_pair(x, d, residual, Ad, t3[1:2]); ev[4].record()
        dn = pre_op(residual); ev[5].record()
        beta = dot_op.fused(dn, Ad, den=t3[0:1], scale=-1.0)[1:2]; ev[6].record()
        d = cd_solve._axpy(dn, beta, d); ev[7].record()
        r2 = dot_op(residual, residual); ev[8].record()
        torch.cuda.synchronize()
        if it > 0:
            tot += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(labels))])
    for l, t in zip(labels, tot / (nit - 1)):
        print('%-28s %8.3f ms' % (l, t), file=sys.stderr)
    print('%-28s %8.3f ms' % ('sum', tot.sum() / (nit - 1)), file=sys.stderr)
