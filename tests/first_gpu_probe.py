"""Diagnostic script for the first GPU contact: prints errors and rough timings instead of asserting."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import rand_alm, rel_l2
from plancklens_b200 import sht
from oracle import ref_sht

def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print(torch.cuda.get_device_name(0))
rng = np.random.default_rng(0)
for nside, lmax in [(16, 40), (64, 128)]:
    plan = sht.get_plan(nside, lmax)
    for spin in range(4):
        if spin == 0:
            a = rand_alm(rng, lmax)
            X, _ = plan.legendre_synth(0, sht.dev_alm(a)); torch.cuda.synchronize()
            R, _ = ref_sht.legendre_synth(nside, 0, lmax, lmax, a)
            print(nside, lmax, spin, 'leg synth', rel_l2(X.cpu().numpy()[:, :lmax+1], R))
            m = plan.ring_synth(X); torch.cuda.synchronize()
            print(nside, lmax, spin, 'ring synth', rel_l2(m.cpu().numpy(), ref_sht.phase2map(nside, R)))
            mm = rng.standard_normal(12*nside**2)
            Y = plan.ring_anal(sht.dev_map(mm)); torch.cuda.synchronize()
            RY = ref_sht.map2phase(nside, mm, lmax) * (4*np.pi/(12*nside**2))
            print(nside, lmax, spin, 'ring anal', rel_l2(Y.cpu().numpy()[:, :lmax+1], RY))
            b, _ = plan.legendre_anal(0, Y); torch.cuda.synchronize()
            print(nside, lmax, spin, 'map2alm', rel_l2(b.cpu().numpy(), ref_sht.map2alm(mm, lmax=lmax)))
        else:
            g, c = rand_alm(rng, lmax, spin), rand_alm(rng, lmax, spin)
            X1, X2 = plan.legendre_synth(spin, sht.dev_alm(g), sht.dev_alm(c)); torch.cuda.synchronize()
            R1, R2 = ref_sht.legendre_synth(nside, spin, lmax, lmax, g, c)
            print(nside, lmax, spin, 'leg synth', rel_l2(X1.cpu().numpy()[:, :lmax+1], R1), rel_l2(X2.cpu().numpy()[:, :lmax+1], R2))
            mm = [rng.standard_normal(12*nside**2) for _ in range(2)]
            ga, ca = plan.map2alm_spin(sht.dev_map(mm[0]), sht.dev_map(mm[1]), spin); torch.cuda.synchronize()
            rg_, rc_ = ref_sht.map2alm_spin(mm, spin, lmax=lmax)
            print(nside, lmax, spin, 'map2alm_spin', rel_l2(ga.cpu().numpy(), rg_), rel_l2(ca.cpu().numpy(), rc_))
# timings at full size
nside, lmax = 2048, 2048
t0 = time.time(); plan = sht.get_plan(nside, lmax); print('plan create s', time.time() - t0)
a = sht.dev_alm(rand_alm(rng, lmax)); c = sht.dev_alm(rand_alm(rng, lmax, 2))
for spin in range(4):
    t0 = time.time()
    if spin == 0: plan.legendre_synth(0, a)
    else: plan.legendre_synth(spin, a, c)
    torch.cuda.synchronize(); print('first call (tables+seeds) spin', spin, time.time() - t0, 's')
X1 = plan.new_phase(); X2 = plan.new_phase()
nlm = sum(lmax - m + 1 for m in range(lmax + 1))
F0 = 8 * nlm * 2 * nside; Fs = 24 * nlm * 2 * nside
for spin in range(4):
    if spin == 0:
        ms = t_ms(lambda: plan.legendre_synth(0, a, X1=X1)); print('leg synth s0 ms', ms, 'TF/s', F0/ms/1e9)
        ms = t_ms(lambda: plan.legendre_anal(0, X1)); print('leg anal  s0 ms', ms, 'TF/s', F0/ms/1e9)
    else:
        ms = t_ms(lambda: plan.legendre_synth(spin, a, c, X1=X1, X2=X2)); print('leg synth s%d ms' % spin, ms, 'TF/s', Fs/ms/1e9)
        ms = t_ms(lambda: plan.legendre_anal(spin, X1, X2)); print('leg anal  s%d ms' % spin, ms, 'TF/s', Fs/ms/1e9)
mp = torch.empty(12*nside**2, dtype=torch.float64, device='cuda')
ms = t_ms(lambda: plan.ring_synth(X1, out=mp)); print('ring synth ms', ms)
ms = t_ms(lambda: plan.ring_anal(mp, X=X1)); print('ring anal ms', ms)
ms = t_ms(lambda: plan.alm2map(a, out=mp)); print('alm2map ms', ms)
print('plan bytes', plan.device_bytes()/1e9, 'GB; launches', sht._lib.launch_count())
