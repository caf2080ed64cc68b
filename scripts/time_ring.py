import os, sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import rand_alm
from plancklens_b200 import sht
def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nside, lmax = 2048, 2048
rng = np.random.default_rng(0)
plan = sht.Plan(nside, lmax)
a = sht.dev_alm(rand_alm(rng, lmax, 2)); c = sht.dev_alm(rand_alm(rng, lmax, 2))
mp = torch.empty(plan.npix, dtype=torch.float64, device='cuda'); m2 = torch.empty_like(mp)
X1, X2 = plan.legendre_synth(2, a, c)
tl = t_ms(lambda: plan.legendre_synth(2, a, c, X1=X1, X2=X2))
print('NB4_MAXM', os.environ.get('PLK_FFT_NB4_MAXM'), 'ring synth (no mtop) %.3f ms' % t_ms(lambda: plan.ring_synth(X1, out=mp)),
      'ring anal %.3f' % t_ms(lambda: plan.ring_anal(mp, X=X1)),
      'alm2map s2 %.3f (leg %.3f)' % (t_ms(lambda: plan.alm2map_spin(a, c, 2, out=(mp, m2))), tl),
      'map2alm s2 %.3f' % t_ms(lambda: plan.map2alm_spin(mp, m2, 2)), 'alm2map s0 %.3f' % t_ms(lambda: plan.alm2map(a, out=mp)))
