"""Rank / size / barrier shim with the reference's names (reference: plancklens/helpers/mpi.py:17-53).

The reference strides jobs over MPI ranks (`jobs[rank::size]`) and only ever uses rank, size and barrier.
Here one process drives one GPU: when launched under torchrun (RANK / WORLD_SIZE in the environment) rank and
size come from there and `barrier` maps to torch.distributed (NCCL on GPUs, gloo on CPU-only hosts);
otherwise the single-rank stubs of the reference apply.
"""
import os

rank = int(os.environ.get('RANK', 0))
size = int(os.environ.get('WORLD_SIZE', 1))
ANY_SOURCE = 0


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()
    return -1


def bcast(obj, root=0):
    d = _dist()
    if d is None:
        return obj
    box = [obj]
    d.broadcast_object_list(box, src=root)
    return box[0]


def allreduce_sum(arr):
    """Sum of a numpy array over ranks (the mean-field reduction of qest.py:239-243 when simulations are sharded).
    NCCL when a GPU is present (the array is staged on the current device), gloo otherwise; identity on one rank."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return arr
    import numpy as np
    import torch
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a.view(np.float64) if np.iscomplexobj(a) else a.astype(np.float64))
    if d.get_backend() == 'nccl':
        t = t.cuda()
    d.all_reduce(t)
    out = t.cpu().numpy()
    return out.view(np.complex128) if np.iscomplexobj(a) else out


def init(backend=None):
    """Joins the process group torchrun described in the environment (no-op on a single process)."""
    global rank, size
    import torch
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', 1)) > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group(backend)
    if dist.is_available() and dist.is_initialized():
        rank, size = dist.get_rank(), dist.get_world_size()
    return rank, size


def send(_, dest):
    return 0


def receive(_, source):
    return 0


def finalize():
    d = _dist()
    if d is not None:
        d.destroy_process_group()
    return -1
