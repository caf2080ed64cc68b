import os, sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import rand_alm
from plancklens_b200 import sht
def t_ms(fn, n=4):
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
r = [t_ms(lambda: plan.legendre_synth(0, a, X1=X1)), t_ms(lambda: plan.legendre_synth(2, a, c, X1=X1, X2=X2)),
     t_ms(lambda: plan.legendre_anal(0, X1)), t_ms(lambda: plan.legendre_anal(2, X1, X2))]
print(os.environ.get('PLK_LIB_PATH', 'default'), 'NRs', os.environ.get('PLK_NR_SYNS'), os.environ.get('PLK_NR_ANAS'), 'synth0 %.3f synth2 %.3f anal0 %.3f anal2 %.3f ms' % tuple(r))
