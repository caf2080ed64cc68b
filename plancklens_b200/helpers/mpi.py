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


def send(_, dest):
    return 0


def receive(_, source):
    return 0


def finalize():
    d = _dist()
    if d is not None:
        d.destroy_process_group()
    return -1
