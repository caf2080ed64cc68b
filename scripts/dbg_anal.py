import os, sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import rand_alm
from plancklens_b200 import sht
def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nside, lmax = 2048, 2048
rng = np.random.default_rng(0)
plan = sht.get_plan(nside, lmax)
a = sht.dev_alm(rand_alm(rng, lmax, 2)); c = sht.dev_alm(rand_alm(rng, lmax, 2))
X1 = plan.new_phase(); X2 = plan.new_phase()
plan.legendre_synth(2, a, c, X1=X1, X2=X2)
for dbg in [0, 1, 2, 3]:
    os.environ['PLK_DBG_ANA'] = str(dbg)
    for nr in [1, 2]:
        os.environ['PLK_NR_ANAS'] = str(nr); os.environ['PLK_NR_ANA0'] = str(2*nr)
        print('dbg', dbg, 'NR', nr, 'anal s2 ms %.3f' % t_ms(lambda: plan.legendre_anal(2, X1, X2)), ' s0 (NR %d) ms %.3f' % (2*nr, t_ms(lambda: plan.legendre_anal(0, X1))))
