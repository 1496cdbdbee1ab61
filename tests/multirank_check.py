"""2-rank GPU check of the fused NVLink all-reduce + Adam kernel (run on a multi-GPU box):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/multirank_check.py

Compares HotPathRunner(fused_collective=True) with the NCCL all-reduce + Adam path: parameters
must agree between the two paths (fp32 tolerance) and be bit-identical across ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fincflow_b200.stack import FincStack, HotPathRunner, LevelSpec


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lv = [LevelSpec(12, 8, 8, 3, (3, 3)), LevelSpec(24, 4, 4, 2, (3, 3))]
    B = 8
    results = {}
    for fused in (True, False):
        torch.manual_seed(0)
        stack = FincStack(lv).to(dev)
        runner = HotPathRunner(stack, B, dev, slots=1, lr=1e-2, process_group=dist.group.WORLD, fused_collective=fused)
        if fused:
            assert runner.fused_collective, getattr(runner, "fused_collective_error", "no symmetric memory")
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for s in runner.slots:
            for li in range(len(lv)):
                s.acts[li][0].normal_(generator=g)
                s.zin[li].normal_(generator=g)
        w0 = stack.flat.detach().clone()
        runner.prepare()
        stack.flat.data.copy_(w0)           # undo the warm-up updates: compare from identical states
        runner.exp_avg.zero_(); runner.exp_avg_sq.zero_(); runner.adam_step.zero_()
        runner._prepare_weights()
        for _ in range(3):
            runner.step(0)
        torch.cuda.synchronize()
        results[fused] = stack.flat.detach().clone()
        gathered = [torch.empty_like(results[fused]) for _ in range(world)]
        dist.all_gather(gathered, results[fused])
        assert all(torch.equal(gathered[0], t) for t in gathered), "parameters differ between ranks"
        assert not torch.equal(results[fused], w0)
        del runner
    err = (results[True] - results[False]).abs().max().item()
    assert err <= 1e-5, err
    if rank == 0:
        print(f"multirank_check OK: world={world}, fused vs NCCL max-abs parameter difference {err:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
