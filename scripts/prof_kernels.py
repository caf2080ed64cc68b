"""Launches each hot kernel once (after a warm-up) at a given size -- the target of ncu captures."""
import argparse, sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import rand_alm
from plancklens_b200 import sht

ap = argparse.ArgumentParser()
ap.add_argument('--nside', type=int, default=2048)
ap.add_argument('--lmax', type=int, default=2048)
ap.add_argument('--spins', type=str, default='0,2')
ap.add_argument('--reps', type=int, default=2)
a = ap.parse_args()
rng = np.random.default_rng(0)
plan = sht.get_plan(a.nside, a.lmax)
g = sht.dev_alm(rand_alm(rng, a.lmax, 2)); c = sht.dev_alm(rand_alm(rng, a.lmax, 2))
X1, X2 = plan.new_phase(), plan.new_phase()
mp = torch.empty(plan.npix, dtype=torch.float64, device='cuda')
for rep in range(a.reps):
    for spin in [int(s) for s in a.spins.split(',')]:
        if spin == 0:
            plan.legendre_synth(0, g, X1=X1); plan.legendre_anal(0, X1)
        else:
            plan.legendre_synth(spin, g, c, X1=X1, X2=X2); plan.legendre_anal(spin, X1, X2)
    plan.ring_synth(X1, out=mp); plan.ring_anal(mp, X=X1)
torch.cuda.synchronize()
print('done')
