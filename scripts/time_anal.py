"""Times the Legendre analysis kernels alone (spin 0 and spin 2, nside = lmax = 2048) and checks them against the
in-tree library's result: A/B of experimental builds (PLK_LIB_PATH=variants/x.so)."""
import os, sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from plancklens_b200 import sht
def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nside, lmax = 2048, int(os.environ.get('PLK_LMAX', 2048))
plan = sht.get_plan(nside, lmax)
g = torch.Generator(device='cuda'); g.manual_seed(1)
X1 = torch.randn((plan.nring, plan.pitch, 2), dtype=torch.float64, device='cuda', generator=g)
X2 = torch.randn((plan.nring, plan.pitch, 2), dtype=torch.float64, device='cuda', generator=g)
X1 = torch.view_as_complex(X1).contiguous(); X2 = torch.view_as_complex(X2).contiguous()
try:
    r0 = t_ms(lambda: plan.legendre_anal(0, X1))
    r2 = t_ms(lambda: plan.legendre_anal(2, X1, X2))
    a = plan.legendre_anal(2, X1, X2)
    chk = float(torch.linalg.norm(a[0]).item())
    print(os.environ.get('PLK_LIB_PATH', 'default'), 'anal0 %.3f anal2 %.3f ms  |G| %.12e' % (r0, r2, chk))
except Exception as ex:
    print(os.environ.get('PLK_LIB_PATH', 'default'), 'failed', ex)
