"""Times the Legendre kernels for the NR (ring pairs per thread) variants selected through PLK_NR_* env vars."""
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

nside = int(os.environ.get('NSIDE', 2048)); lmax = int(os.environ.get('LMAX', 2048))
rng = np.random.default_rng(0)
plan = sht.get_plan(nside, lmax)
a = sht.dev_alm(rand_alm(rng, lmax, 2)); c = sht.dev_alm(rand_alm(rng, lmax, 2))
X1 = plan.new_phase(); X2 = plan.new_phase()
nlm = sum(lmax - m + 1 for m in range(lmax + 1))
F0 = 8 * nlm * 2 * nside; Fs = 24 * nlm * 2 * nside
for nr in [1, 2, 4]:
    for k in ['PLK_NR_SYN0', 'PLK_NR_SYNS', 'PLK_NR_ANA0', 'PLK_NR_ANAS']: os.environ[k] = str(nr)
    ms = t_ms(lambda: plan.legendre_synth(0, a, X1=X1)); print('NR', nr, 'synth s0 ms %.3f TF/s %.2f' % (ms, F0/ms/1e9))
    ms = t_ms(lambda: plan.legendre_anal(0, X1)); print('NR', nr, 'anal  s0 ms %.3f TF/s %.2f' % (ms, F0/ms/1e9))
    ms = t_ms(lambda: plan.legendre_synth(2, a, c, X1=X1, X2=X2)); print('NR', nr, 'synth s2 ms %.3f TF/s %.2f' % (ms, Fs/ms/1e9))
    ms = t_ms(lambda: plan.legendre_anal(2, X1, X2)); print('NR', nr, 'anal  s2 ms %.3f TF/s %.2f' % (ms, Fs/ms/1e9))
mp = torch.empty(12*nside**2, dtype=torch.float64, device='cuda')
ms = t_ms(lambda: plan.ring_synth(X1, out=mp)); print('ring synth (no mtop) ms %.3f' % ms)
ms = t_ms(lambda: plan.ring_anal(mp, X=X1)); print('ring anal (no mtop) ms %.3f' % ms)
for k in ['PLK_NR_SYN0', 'PLK_NR_SYNS', 'PLK_NR_ANA0', 'PLK_NR_ANAS']: os.environ.pop(k)
ms = t_ms(lambda: plan.alm2map(a, out=mp)); print('alm2map s0 ms %.3f' % ms)
ms = t_ms(lambda: plan.map2alm(mp)); print('map2alm s0 ms %.3f' % ms)
m2 = torch.empty_like(mp)
ms = t_ms(lambda: plan.alm2map_spin(a, c, 2, out=(mp, m2))); print('alm2map s2 ms %.3f' % ms)
ms = t_ms(lambda: plan.map2alm_spin(mp, m2, 2)); print('map2alm s2 ms %.3f' % ms)
