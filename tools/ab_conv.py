"""Timing of the forward / backward-input / backward-weight kernels as CUDA graphs of N launches
(L2 flushed for large batches).  python tools/ab_conv.py [--batches 256,16384] [--k 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="256,16384")
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--shapes", default="12x16x16,24x8x8,48x4x4")
ap.add_argument("--chain", type=int, default=8)
ap.add_argument("--kernels", default="forward,backward_input,backward_weight")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ws = _native.new_workspace(256 << 20, dev)
for shp in args.shapes.split(","):
    CT, H, W = (int(v) for v in shp.split("x"))
    k = args.k
    w = FastFlowUnit(CT, CT, (k, k)).weight.detach().to(dev)
    for B in (int(b) for b in args.batches.split(",")):
        x = torch.randn(B, CT, H, W, device=dev); dz = torch.randn_like(x); y = torch.empty_like(x); dw = torch.empty_like(w)
        fns = {"forward": lambda: _native.forward(x, w, out=y, want_logdet=False),
               "backward_input": lambda: _native.backward_input(dz, w, out=y),
               "backward_weight": lambda: _native.backward_weight(dz, x, (k, k), out=dw, workspace=ws)}
        for name in args.kernels.split(","):
            fn = fns[name]
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(args.chain):
                    fn()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / args.chain)
            ts.sort()
            us = ts[2]
            print(f"{name:16s} [{B},{CT},{H},{W}] k={k}: {us:8.2f} us/launch  {8.0 * B * CT * H * W / us / 1e3:7.1f} GB/s  "
                  f"{2.0 * B * H * W * CT * (CT // 4) * k * k / us / 1e6:6.2f} TFLOP/s", flush=True)
