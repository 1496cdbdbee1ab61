"""phase times of the headline workload for the runner's execution options, results compared bit for bit
(run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fincflow_b200.stack import FincStack, HotPathRunner

dev = torch.device("cuda:0")
res = {}
CASES = [("per-unit", dict(chain=False)), ("chain", dict(chain=True)), ("chain+level_parallel", dict(chain=True, level_parallel=True))]
for name, kw in CASES:
    torch.manual_seed(0)
    stack = FincStack(bench.levels()).to(dev)
    r = HotPathRunner(stack, 256, dev, slots=1, **kw)
    g = torch.Generator(device=dev).manual_seed(1)
    for li in range(len(stack.levels)):
        r.slots[0].acts[li][0].normal_(generator=g)
        r.slots[0].zin[li].normal_(generator=g)
    r.prepare()
    for _ in range(5):
        for p in range(4):
            r.run_phase(0, p)
    torch.cuda.synchronize()
    K = 100
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    for i in range(K):
        for p in range(4):
            ev[i][p].record()
            r.run_phase(0, p)
        ev[i][4].record()
    torch.cuda.synchronize()
    ph = [sum(ev[i][p].elapsed_time(ev[i][p + 1]) for i in range(K)) / K for p in range(4)]
    print(f"{name}: " + ", ".join(f"{n} {v * 1e3:.1f} us" for n, v in zip(HotPathRunner.PHASES, ph)),
          f"| step {ev[0][0].elapsed_time(ev[K - 1][4]) / K * 1e3:.1f} us, launches/step {r.launches_per_step}")
    res[name] = (stack.flat.detach().clone(), [t.clone() for t in r.slots[0].logp],
                 [r.slots[0].sample_out[k].clone() for k in sorted(r.slots[0].sample_out)])
a = res["per-unit"]
for name, b in res.items():
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    print(name, "vs per-unit: params rel", f"{rel(b[0], a[0]):.1e}", " logp rel", f"{max(rel(x, y) for x, y in zip(b[1], a[1])):.1e}",
          " samples rel", f"{max(rel(x, y) for x, y in zip(b[2], a[2])):.1e}",
          "(chains are bit-identical; the batched dW sums its batch slices in another fixed order)")
