"""Data-parallel plumbing for the FInC hot path: one process per GPU, batch sharded on dim 0.

The reference scales with single-process `nn.DataParallel` (scatter / replicate / gather,
gradients reduced onto GPU 0: fastflow/fastflow_cifar_multi_gpu.py:439-440).  Here every rank
owns a contiguous slice of the batch and a full replica of the weights:

  * training: the masked FInC weight gradients of all units are written by the wgrad kernel
    straight into ONE flat bucket; the only collective of a step is a single NCCL
    all-reduce(SUM) of that bucket over NVLink/NVSwitch.  With the per-rank loss scaled by
    1/(B_local * world) the sum IS the gradient of the global-batch mean loss
    (reference loss: `sum / len(x)`, train/experiment.py:204).
  * sampling and likelihood evaluation: no collective at all (images are independent).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def dp_env():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n: int, rank: int, world: int):
    """contiguous near-equal slice [start, stop) of n items for `rank`"""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def loss_scale(local_batch: int, world: int) -> float:
    """d(-mean_n logp)/dlogp for the GLOBAL batch, applied to a rank's local samples"""
    return 1.0 / (local_batch * world)


def init_process_group(backend=None, device=None):
    rank, world, local = dp_env()
    if world == 1:
        return None
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return dist.group.WORLD


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """the one collective of a training step: SUM of the flat gradient bucket"""
    if group is not None or (dist.is_initialized() and dist.get_world_size() > 1):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_params_(flat: torch.Tensor, group=None, src: int = 0) -> torch.Tensor:
    """one-time replication of the weights (reference: DataParallel re-broadcasts every step)"""
    if group is not None or (dist.is_initialized() and dist.get_world_size() > 1):
        dist.broadcast(flat, src=src, group=group)
    return flat
