"""Run each hot-path kernel a few times at the bench shapes (eager, no graphs) -- the command
profiled under ncu (see profiles/README.md).  Also prints CUDA-event timings and a
launch-floor measurement (chain of trivial kernels in a CUDA graph)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--floor", action="store_true")
ap.add_argument("--shapes", default="12x16x16,24x8x8,48x4x4")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
for shp in args.shapes.split(","):
    CT, H, W = (int(v) for v in shp.split("x"))
    B, k = args.batch, args.k
    unit = FastFlowUnit(CT, CT, (k, k)).to(dev)
    w = unit.weight.detach()
    x = torch.randn(B, CT, H, W, device=dev)
    dz = torch.randn_like(x)
    y = torch.empty_like(x)
    dw = torch.empty_like(w)
    ws = _native.new_workspace(_native.backward_weight_workspace_bytes(B, 4, CT // 4, H, W, k, k), dev)
    fns = {
        "forward": lambda: _native.forward(x, w, out=y, want_logdet=False),
        "backward_input": lambda: _native.backward_input(dz, w, out=y),
        "backward_weight": lambda: _native.backward_weight(dz, x, (k, k), out=dw, workspace=ws),
        "inverse": lambda: _native.inverse(x, w, out=y),
    }
    for name, fn in fns.items():
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name:16s} [{B},{CT},{H},{W}] k={k}: " + " ".join(f"{t:8.1f}" for t in ts) + " us")

if args.floor:
    t = torch.zeros(32, device=dev)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            t.add_(1.0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(48):
            t.add_(1.0)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    e1.synchronize()
    print(f"launch floor: {e0.elapsed_time(e1) * 1e3 / 20 / 48:.2f} us per dependent trivial kernel in a CUDA graph")
