"""CPU tests (gloo, world_size 2) of the data-parallel host logic: batch sharding and the
flat-bucket all-reduce reproduce the global-batch gradient of the reference's CPU path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err


def test_shard_range_partitions():
    from fincflow_b200.distributed import shard_range

    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fincflow_b200 import distributed as fd
    from fincflow_b200.stack import LevelSpec
    from oracle.reference_path import ReferenceCpuStack

    torch.set_num_threads(1)
    pg = fd.init_process_group("gloo")
    assert fd.dp_env() == (rank, world, rank)
    lv = [LevelSpec(12, 6, 6, 2, (3, 3))]
    B = 6
    full = ReferenceCpuStack(lv, B, seed=3, lr=0.0, threads=1)       # identical on every rank
    lo, hi = fd.shard_range(B, rank, world)
    # local loss = -sum(logp_local) * loss_scale  ->  SUM over ranks == global mean loss gradient
    flat = []
    for u, w in enumerate(full.weights[0]):
        w.grad = None
    h = full.x[0][lo:hi]
    from oracle.reference_path import unit_forward
    import math

    for w in full.weights[0]:
        h = unit_forward(h, w)
    logp = -0.5 * h.flatten(1).pow(2).sum(1) - 0.5 * lv[0].dim * math.log(2 * math.pi)
    (-(logp.sum()) * fd.loss_scale(hi - lo, world) * ((hi - lo) * world / B)).backward()
    bucket = torch.cat([(w.grad * full.masks[0]).flatten() for w in full.weights[0]])
    fd.allreduce_sum_(bucket, pg)
    if rank == 0:
        out.put(bucket.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_global_gradient():
    from fincflow_b200.stack import LevelSpec
    from oracle.reference_path import ReferenceCpuStack, unit_forward
    import math

    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lv = [LevelSpec(12, 6, 6, 2, (3, 3))]
    B = 6
    full = ReferenceCpuStack(lv, B, seed=3, lr=0.0, threads=1)
    h = full.x[0]
    for w in full.weights[0]:
        h = unit_forward(h, w)
    logp = -0.5 * h.flatten(1).pow(2).sum(1) - 0.5 * lv[0].dim * math.log(2 * math.pi)
    (-logp.sum() / B).backward()
    want = torch.cat([(w.grad * full.masks[0]).flatten() for w in full.weights[0]]).numpy()
    assert rel_err(got, want) <= 1e-5


def test_compat_shims_have_the_pybind_signature():
    import inspect

    from fincflow_b200 import compat

    for obj in (compat.cinc_cuda_level1, compat.cinc_cuda_level2):
        assert list(inspect.signature(obj.inverse).parameters) == ["input", "kernel", "output"]


def test_plan_host_cores():
    from fincflow_b200.distributed import _parse_cpulist, plan_host_cores

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    allowed = range(16)
    # no topology: even split
    assert plan_host_cores(3, 8, allowed) == [6, 7]
    # two NUMA nodes, four GPUs each: every rank gets CPUs of ITS node, disjoint from its peers
    nodes = [0, 0, 0, 0, 1, 1, 1, 1]
    node_cpus = {0: list(range(0, 8)), 1: list(range(8, 16))}
    got = [plan_host_cores(r, 8, allowed, nodes, node_cpus) for r in range(8)]
    assert got[0] == [0, 1] and got[3] == [6, 7] and got[4] == [8, 9] and got[7] == [14, 15]
    assert len({c for g in got for c in g}) == 16
    # more ranks than CPUs on a node: fall back to the even split of the allowed set
    assert plan_host_cores(1, 4, range(4), [0, 0, 0, 0], {0: [0, 1]}) == [1]


# ---- whole-flow data-parallel trainer (fincflow_b200/train.py) on a CPU toy flow ----------------------------
class _ToyFlow(torch.nn.Module):
    """ActNorm (data-dependent init) + an elementwise affine 'flow': forward(x) -> (z, logp[B])"""

    def __init__(self):
        super().__init__()
        from fincflow_b200.flows import ActNorm

        self.act = ActNorm(3)
        self.log_a = torch.nn.Parameter(torch.tensor([0.1, -0.2, 0.3]))
        self.b = torch.nn.Parameter(torch.zeros(3))
        self.unused = torch.nn.Parameter(torch.ones(2))      # never receives a gradient

    def forward(self, x):
        h, ld = self.act(x)
        z = h * torch.exp(self.log_a).view(1, 3, 1, 1) + self.b.view(1, 3, 1, 1)
        ld = ld + self.log_a.sum() * x.shape[2] * x.shape[3]
        return z, -0.5 * z.flatten(1).pow(2).sum(1) + ld


def _toy_batch():
    g = torch.Generator().manual_seed(5)
    return torch.randn(8, 3, 4, 4, generator=g) * 2.0 + 1.0


def _trainer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fincflow_b200.train import FlowTrainer

    torch.set_num_threads(1)
    dist.init_process_group("gloo")
    torch.manual_seed(100 + rank)                              # replicas start DIFFERENT: the trainer must sync them
    model = _ToyFlow()
    with torch.no_grad():
        model.b.add_(torch.randn(3))
    trainer = FlowTrainer(model, lr=1e-2, bucket_mb=1e-5)      # tiny buckets: several all-reduces per step
    assert len(trainer.buckets) >= 3
    x = _toy_batch()
    shard = x[rank * 4:(rank + 1) * 4]
    losses = [float(trainer.step(shard)) for _ in range(3)]
    diff = trainer.replica_max_diff()
    if rank == 0:
        out.put((losses, diff, {k: v.numpy().copy() for k, v in model.state_dict().items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_flow_trainer_two_ranks_equals_single_process():
    from fincflow_b200.train import FlowTrainer

    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    losses, diff, sd = out.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert diff == 0.0                                          # replicas bit-identical after three steps
    # single process on the full batch, starting from rank 0's initial parameters
    torch.manual_seed(100)
    model = _ToyFlow()
    with torch.no_grad():
        model.b.add_(torch.randn(3))
    trainer = FlowTrainer(model, lr=1e-2)
    x = _toy_batch()
    want = [float(trainer.step(x)) for _ in range(3)]
    assert np.allclose(losses, want, rtol=1e-5)
    for k, v in model.state_dict().items():
        assert np.allclose(sd[k], v.numpy(), rtol=1e-5, atol=1e-6), k   # incl. ActNorm initialised from the GLOBAL batch
