"""CPU, world_size 2 over gloo: the N > 1 host logic -- rank/size/barrier shim (joined lazily at the first barrier),
simulation striding and the mean-field all-reduce of qest.library.get_sim_qlm_mf_sharded, and the reference-shaped
get_sim_qlm_mf staying a LOCAL call (no GPU involved: get_sim_qlm is stubbed)."""
import os
import subprocess
import sys
import tempfile


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
from plancklens_b200.helpers import mpi
os.environ['CUDA_VISIBLE_DEVICES'] = ''         # CPU test: the lazy join must pick gloo
assert mpi.size == 2 and mpi.rank == int(os.environ['RANK'])      # from torchrun's environment, before any join
mpi.barrier()                                   # no explicit mpi.init(): the first barrier joins the process group
import torch.distributed as dist
assert dist.is_initialized() and dist.get_world_size() == 2
rank, size = mpi.rank, mpi.size
from plancklens_b200 import qest, hp
lib = object.__new__(qest.library)
lib.lib_dir = %(tmp)r
lib.lmax_qlm = {'T': 8, 'P': 8, 'PS': 8}
lib.keys_fund = ['ptt']; lib.keys_remaps = {}
calls = []
def fake(k, idx, lmax=None):
    calls.append(int(idx))
    return (idx + 1) * (np.arange(hp.Alm.getsize(8)) + 1j)
lib.get_sim_qlm = fake
mf = lib.get_sim_qlm_mf_sharded('ptt', np.arange(6))
expect = np.mean([i + 1 for i in range(6)]) * (np.arange(hp.Alm.getsize(8)) + 1j)
assert np.allclose(mf, expect), (mf[:3], expect[:3])
assert calls == list(range(6))[rank::2], calls          # each rank evaluated only its share
# the reference-shaped call is local: ranks may call it with different arguments, or not at all, without pairing up
del calls[:]
if rank == 1:
    mf1 = lib.get_sim_qlm_mf('ptt', np.array([7, 9]))
    assert np.allclose(mf1, 9.0 * (np.arange(hp.Alm.getsize(8)) + 1j))
    assert calls == [7, 9], calls
assert np.allclose(mpi.allreduce_sum(np.array([1.0 + rank])), 3.0)
assert mpi.bcast('x' if rank == 0 else None) == 'x'
mpi.barrier()
open(os.path.join(%(tmp)r, 'ok_%%d' %% rank), 'w').write('ok')
mpi.finalize()
'''


def test_mean_field_sharded_over_two_ranks():
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, 'w.py')
        with open(script, 'w') as f:
            f.write(WORKER % {'root': ROOT, 'tmp': tmp})
        env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29533')
        out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                              '--master-addr', '127.0.0.1', '--master-port', '29533', script],
                             capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert os.path.exists(os.path.join(tmp, 'ok_0')) and os.path.exists(os.path.join(tmp, 'ok_1'))
