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


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def plan_host_cores(local_rank: int, local_world: int, allowed, gpu_numa_nodes=None, node_cpus=None):
    """CPU set for one rank of a one-process-per-GPU job: the ranks whose GPU sits on the same NUMA
    node split that node's allowed CPUs evenly; without topology information the allowed CPUs are
    split evenly over all ranks.  Pure function (unit-tested on CPU)."""
    allowed = sorted(allowed)
    if gpu_numa_nodes is not None and node_cpus is not None and gpu_numa_nodes[local_rank] in node_cpus:
        node = gpu_numa_nodes[local_rank]
        peers = [r for r in range(local_world) if gpu_numa_nodes[r] == node]
        pool = [c for c in node_cpus[node] if c in set(allowed)]
        if len(pool) >= len(peers):
            k = peers.index(local_rank)
            per = len(pool) // len(peers)
            return pool[k * per:(k + 1) * per]
    if len(allowed) >= local_world:
        per = len(allowed) // local_world
        return allowed[local_rank * per:(local_rank + 1) * per]
    return allowed


def bind_host_cores(local_rank: int, local_world: int):
    """Pin this process (and the pinned host buffers it allocates afterwards, by first touch) to the
    CPUs next to its GPU.  With 8 ranks feeding 8 GPUs from one host the host->device copies of the
    end-to-end path otherwise cross sockets and the ranks' launch threads migrate.  Best effort:
    returns the CPU list, or None when the platform gives no affinity control."""
    if local_world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = os.sched_getaffinity(0)
        nodes, node_cpus = None, None
        try:
            nodes = []
            for r in range(local_world):
                p = torch.cuda.get_device_properties(r)
                bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
                with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
                    nodes.append(int(f.read()))
            node_cpus = {}
            for n in set(nodes):
                if n >= 0:
                    with open(f"/sys/devices/system/node/node{n}/cpulist") as f:
                        node_cpus[n] = _parse_cpulist(f.read())
        except Exception:
            nodes, node_cpus = None, None
        cpus = plan_host_cores(local_rank, local_world, allowed, nodes, node_cpus)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
