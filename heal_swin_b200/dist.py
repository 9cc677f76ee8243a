"""Data-parallel plumbing for the hot path: one process per GPU, the batch sharded over ranks,
identical weights, no data-path collective; the only exchange step is the gradient all-reduce (the reference gets it
from Lightning's DDPPlugin, heal_swin/train.py:187).  ``torch.distributed`` does the transport (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests); nothing here touches the kernels.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process -> (0, 0, 1))."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend="nccl", device=None):
    """Initialise the default process group when launched with WORLD_SIZE > 1; returns (rank, local_rank, world)."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def shard_range(global_batch, rank, world):
    """Contiguous [lo, hi) slice of the global batch owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(tensors):
    """In-place mean over ranks of a list of tensors (what DDP does to the gradients), one flat bucket."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def max_over_ranks(value, device="cpu"):
    """Max of a python float over ranks (device timings are reported as the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def wrap_ddp(model, local_rank):
    """torch DDP over the ordinary nn.Parameters of the drop-in modules (bucketed, overlapped with backward)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True)
