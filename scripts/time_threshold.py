"""Start threshold of the Legendre recurrences (default 2^-60, libsharp's sharp_ftol; 2^-120 until late in round 2): walked share of the
(l, m, ring pair) volume, kernel-stage times and the change of the results, nside = lmax = 2048 (or argv[1], argv[2])."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plancklens_b200 import sht

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(0)
n = sht.alm_size(lmax)
ls = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
red = 1.0 / (1.0 + ls) ** 2          # red spectrum: the dropped low-l terms carry the largest coefficients
g = sht.dev_alm((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * red * (ls >= 2))
c = sht.dev_alm((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * red * (ls >= 2))
t = sht.dev_alm((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * red)
m1 = sht.dev_map(rng.standard_normal(12 * nside ** 2))
m2 = sht.dev_map(rng.standard_normal(12 * nside ** 2))
plan = sht.get_plan(nside, lmax)


def timed(f, nrep=5):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        r = f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep, r


ref = None
for thr in (-200, -120, -90, -60, -45):
    plan.set_seed_threshold(thr)
    t0, a = timed(lambda: plan.alm2map(t))
    t2, b = timed(lambda: plan.alm2map_spin(g, c, 2))
    t2a, d = timed(lambda: plan.map2alm_spin(m1, m2, 2))
    t0a, e = timed(lambda: plan.map2alm(m1))
    res = [a, b[0], b[1], d[0], d[1], e]
    if ref is None:
        ref = [x.clone() for x in res]
    err = [float(torch.linalg.norm(x - y) / torch.linalg.norm(y)) for x, y in zip(res, ref)]
    errmax = [float((x - y).abs().max() / y.abs().max()) for x, y in zip(res, ref)]
    print('2^%d: share spin0 %.4f spin2 %.4f | alm2map %.3f ms, alm2map_spin2 %.3f, map2alm_spin2 %.3f, map2alm %.3f | rel L2 vs 2^-200: %s | max-abs/max: %s' %
          (thr, plan.active_fraction(0), plan.active_fraction(2), t0, t2, t2a, t0a, ' '.join('%.1e' % x for x in err), ' '.join('%.1e' % x for x in errmax)))
