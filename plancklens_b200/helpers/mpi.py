"""Rank / size / barrier shim with the reference's names (reference: plancklens/helpers/mpi.py:17-53).

The reference strides jobs over MPI ranks (`jobs[rank::size]`) and only ever uses rank, size and barrier.
Here one process drives one GPU: when launched under torchrun (RANK / WORLD_SIZE in the environment) rank and
size come from there and `barrier` maps to torch.distributed (NCCL on GPUs, gloo on CPU-only hosts);
otherwise the single-rank stubs of the reference apply.

rank, size and barrier always go together, as in the reference (mpi.py:30-36): with WORLD_SIZE > 1 the first
`barrier` / `bcast` / `allreduce_sum` JOINS the process group torchrun described (lazily, so that importing this
module stays free), and raises if that is impossible -- a rank != 0 never runs past a dummy barrier into files rank 0
is still writing (`if mpi.rank == 0: ...; mpi.barrier()` blocks of qest, qecl, filt_*, nhl, qresp, n1, sql).
"""
import os

rank = int(os.environ.get('RANK', 0))
size = int(os.environ.get('WORLD_SIZE', 1))
ANY_SOURCE = 0


def _dist():
    """torch.distributed once a process group exists; joins it on first use when launched with WORLD_SIZE > 1"""
    import torch.distributed as dist
    if not dist.is_available():
        if size > 1:
            raise RuntimeError("WORLD_SIZE = %d but torch.distributed is unavailable: no barrier exists" % size)
        return None
    if not dist.is_initialized():
        if size <= 1:
            return None
        init()          # raises if the rendezvous variables are missing
    return dist


def require_group():
    """Raises unless collectives are real: single process, or a joined process group whose size matches `size`."""
    d = _dist()
    if size > 1:
        if d is None or d.get_world_size() != size:
            raise RuntimeError("mpi.size = %d but the process group has %s ranks"
                               % (size, 'no' if d is None else d.get_world_size()))


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
    """Sum of a numpy array over ranks (the sharded mean-field reduction, qest.library.get_sim_qlm_mf_sharded).
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
        for var in ('MASTER_ADDR', 'MASTER_PORT', 'RANK'):
            if var not in os.environ:
                raise RuntimeError("WORLD_SIZE = %s but %s is not set: cannot join the process group (launch with "
                                   "torchrun, or unset WORLD_SIZE for a single process)" % (os.environ['WORLD_SIZE'], var))
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
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
    return -1
