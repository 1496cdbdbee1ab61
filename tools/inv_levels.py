"""per-level timing of the wavefront inverse at the ImageNet64 unit shapes (run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

dev = torch.device("cuda:0")
for B in (1024, 2048):
    for k in (3, 5):
        for (C, H, W) in ((12, 32, 32), (24, 16, 16), (48, 8, 8), (96, 4, 4)):
            torch.manual_seed(0)
            unit = FastFlowUnit(C, C, (k, k)).to(dev)
            z = torch.randn(B, C, H, W, device=dev)
            x = torch.empty_like(z)
            w = unit.weight.detach()
            for _ in range(3):
                _native.inverse(z, w, out=x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                _native.inverse(z, w, out=x)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 10 * 1e3
            cq = C // 4
            flops = 2.0 * B * H * W * C * cq * k * k
            line = f"B={B} k={k} [{C},{H},{W}] Cq={cq}: {us:8.1f} us  {flops / us / 1e6:6.2f} TFLOP/s  {8 * z.numel() / us / 1e3:7.1f} GB/s"
            if _native.inverse_dense_bytes(4, cq, H, W) > 0:
                blob = _native.inverse_dense_prepare(w, H, W)
                xd = torch.empty_like(z)
                for _ in range(3):
                    _native.inverse_dense(z, blob, out=xd)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    _native.inverse_dense(z, blob, out=xd)
                e1.record()
                torch.cuda.synchronize()
                usd = e0.elapsed_time(e1) / 10 * 1e3
                err = float((xd - x).abs().max() / x.abs().max())
                zz, _ = unit(xd)
                line += f" | dense GEMM {usd:7.1f} us, vs wavefront rel {err:.2e}, round trip {float((zz - z).abs().max()):.2e}"
            print(line)
