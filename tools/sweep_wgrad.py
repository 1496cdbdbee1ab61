"""Empirical sweep of the wgrad plan space (FINC_WG_FORCE, read at every launch) against the
cost model's choice; one process, one CUDA graph of 8 launches per candidate.
python tools/sweep_wgrad.py [--flags 128] [--batch 256] [--shapes 12x16x16,...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

ap = argparse.ArgumentParser()
ap.add_argument("--flags", type=int, default=128)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--shapes", default="12x16x16,24x8x8,48x4x4")
ap.add_argument("--top", type=int, default=6)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
ws = _native.new_workspace(256 << 20, dev)
B, flags = args.batch, args.flags


def timed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 8)
    ts.sort()
    return ts[2]


for shp in args.shapes.split(","):
    CT, H, W = (int(v) for v in shp.split("x"))
    C = CT // 4
    w = FastFlowUnit(CT, CT, (3, 3)).weight.detach().to(dev)
    x = torch.randn(B, CT, H, W, device=dev); dz = torch.randn_like(x); dw = torch.empty_like(w)
    fn = lambda: _native.backward_weight(dz, x, (3, 3), out=dw, workspace=ws, flags=flags)
    os.environ.pop("FINC_WG_FORCE", None)
    ref = dw.clone()
    fn(); torch.cuda.synchronize(); ref.copy_(dw)
    res = [(timed(fn), (0, 0, 0))]
    for ob in (1, 2, 3, 4, 6):
        if ob > C:
            continue
        nob = -(-C // ob)
        for z in sorted({z for z in (1, 2, 3, 4, 6, 9, 12, 18, 24) if z <= C * nob}):
            for ch in (0, 16, 8, 4):
                os.environ["FINC_WG_FORCE"] = f"{ob},{z},{ch}"
                try:
                    t = timed(fn)
                except Exception:
                    continue
                ok = torch.allclose(dw, ref, rtol=1e-4, atol=1e-4 * ref.abs().max().item())
                res.append((t, (ob, z, ch) if ok else ("BAD", ob, z, ch)))
    os.environ.pop("FINC_WG_FORCE", None)
    print(f"=== [{B},{CT},{H},{W}] flags={flags}: model choice {res[0][0]:.2f} us", flush=True)
    for t, force in sorted(res, key=lambda r: r[0])[:args.top]:
        print(f"  {t:7.2f} us  force={force}", flush=True)
