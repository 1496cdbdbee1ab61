"""A/B timing of the inverse kernel generations (register-window vs shared-memory wavefront):
each variant as a CUDA graph of N dependent launches ping-ponging between two buffers,
L2 flushed between replays for the large batches.  python tools/ab_inverse.py [--batches 256,16384]"""
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
ap.add_argument("--chain", type=int, default=16)
ap.add_argument("--variants", default="rw,wave")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
FLAGS = {"rw": 0, "wave": _native.FLAG_WAVE_SMEM, "generic": _native.FLAG_GENERIC_TILED}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for shp in args.shapes.split(","):
    CT, H, W = (int(v) for v in shp.split("x"))
    k = args.k
    w = FastFlowUnit(CT, CT, (k, k)).weight.detach().to(dev)
    for B in (int(b) for b in args.batches.split(",")):
        bufs = [torch.randn(B, CT, H, W, device=dev) * 0.1 for _ in range(2)]
        nb = _native.prepared_weights_bytes(_native.PREP_INVERSE, B, 4, CT // 4, H, W, k, k)
        table = None
        if nb:
            table = torch.empty((1, nb), dtype=torch.uint8, device=dev)
            _native.prepare_weights(w[None], table, _native.PREP_INVERSE, B, H, W)
        for var in args.variants.split(","):
            def chain():
                for i in range(args.chain):
                    if table is not None:
                        _native.inverse(bufs[i & 1], None, out=bufs[(i + 1) & 1], flags=FLAGS[var], prepared=table[0], ksize=(k, k))
                    else:
                        _native.inverse(bufs[i & 1], w, out=bufs[(i + 1) & 1], flags=FLAGS[var])
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                chain()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                chain()
            ts = []
            for _ in range(5):
                for b in bufs:
                    b.normal_().mul_(0.1)
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / args.chain)
            ts.sort()
            us = ts[len(ts) // 2]
            gb = 8.0 * B * CT * H * W / us / 1e3
            fl = 2.0 * B * H * W * CT * (CT // 4) * k * k / us / 1e6
            print(f"inverse {var:7s} [{B},{CT},{H},{W}] k={k}: {us:8.2f} us/launch  {gb:7.1f} GB/s  {fl:6.2f} TFLOP/s", flush=True)
