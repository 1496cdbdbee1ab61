"""On-device check of the tensor-core path (run under gpurun): every case in its own
subprocess with a timeout, so a trap in one kernel does not hide the others.

    python tools/tc_check.py            # all cases
    python tools/tc_check.py --case conv:2,16,16,64,256,1,3
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = [
    # conv: B,H,W,Cin,N,taps,npass
    "conv:2,16,16,32,32,1,1", "conv:2,16,16,32,32,1,3", "conv:2,16,16,64,256,1,1", "conv:2,16,16,64,256,1,3",
    "conv:2,16,16,64,32,9,3", "conv:3,14,14,32,64,9,3", "conv:5,7,7,96,128,9,3", "conv:9,4,4,64,64,9,3",
    "conv:2,32,32,32,32,9,3", "conv:8,16,16,512,512,1,3", "conv:8,16,16,512,512,1,1", "conv:300,8,8,512,512,1,3",
    # wgrad: np,M,N,npass   dW[m,n] = sum_p P[p,m] Q[p,n]
    "wgrad:4096,128,64,3", "wgrad:4096,512,128,3", "wgrad:5000,512,160,3", "wgrad:65536,512,512,3", "wgrad:65536,512,512,1",
    "wgrad:65536,512,64,3", "wgrad:65536,512,160,3",
    # coupling: B,C,H,W,width,npass
    "coupling:4,12,16,16,512,3", "coupling:4,12,16,16,512,1", "coupling:6,24,8,8,512,3", "coupling:9,48,4,4,512,3",
    "coupling:3,4,14,14,64,3", "coupling:3,8,7,7,64,3", "coupling:2,96,4,4,512,3", "coupling:2,12,32,32,512,3",
    # timing at the CIFAR-10 shapes (cfg3, batch 256)
    "time:256,12,16,16,512,3", "time:256,12,16,16,512,1", "time:256,24,8,8,512,3", "time:256,48,4,4,512,3",
]


def run_case(spec):
    import torch
    import torch.nn.functional as F

    from fincflow_b200 import _native

    kind, args = spec.split(":")
    a = [int(v) for v in args.split(",")]
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    if kind == "conv":
        B, H, W, Cin, N, taps, npass = a
        k = 3 if taps == 9 else 1
        x = torch.randn(B, Cin, H, W, device=dev)
        w = torch.randn(N, Cin, k, k, device=dev) / (Cin * taps) ** 0.5
        bias = torch.randn(N, device=dev)
        wp = _native.tc_conv_prepare_weights(w, 0)
        xn = x.permute(0, 2, 3, 1).contiguous()
        flags = _native.FLAG_TF32_1PASS if npass == 1 else 0
        y = _native.tc_conv_nhwc(xn, wp, bias, N, taps, relu=True, flags=flags)
        torch.cuda.synchronize()
        ref = F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)).permute(0, 2, 3, 1)
        err = float((y.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = False
        y32 = F.relu(F.conv2d(x, w, bias, padding=k // 2)).permute(0, 2, 3, 1)
        err32 = float((y32.double() - ref).abs().max() / ref.abs().max())
        print(f"   (cuDNN fp32 vs fp64: {err32:.3e}; ours vs cuDNN fp32: {float((y - y32).abs().max() / ref.abs().max()):.3e})")
        bad = int(((y.double() - ref).abs() > 1e-2 * ref.abs().max()).sum())
        print(f"{spec}: max-norm rel err {err:.3e}, elements off by >1%: {bad} of {ref.numel()}")
        if bad:
            d = (y.double() - ref).abs() > 1e-2 * ref.abs().max()
            idx = d.nonzero()[:8].tolist()
            print("   first bad (b,h,w,n):", idx)
            print("   bad per n%32:", d.sum(dim=(0, 1, 2)).view(-1, 32).sum(0).tolist() if N % 32 == 0 else "")
            print("   bad per (h,w):", d.sum(dim=(0, 3)).tolist() if H * W <= 64 else d.sum(dim=(0, 2, 3)).tolist())
        return err < (2e-3 if npass == 1 else 2e-6)
    if kind == "wgrad":
        npix, M, N, npass = a
        P = torch.randn(npix, M, device=dev)
        Q = torch.randn(npix, N, device=dev)
        flags = _native.FLAG_TF32_1PASS if npass == 1 else 0
        dW = _native.tc_wgrad(P, Q, flags=flags)
        torch.cuda.synchronize()
        ref = P.double().t() @ Q.double()
        err = float((dW.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cuda.matmul.allow_tf32 = False
        r32 = P.t() @ Q
        err32 = float((r32.double() - ref).abs().max() / ref.abs().max())
        bad = int(((dW.double() - ref).abs() > 1e-2 * ref.abs().max()).sum())
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            _native.tc_wgrad(P, Q, flags=flags, out=dW)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n):
            torch.matmul(P.t(), Q, out=r32)
        e1.record()
        torch.cuda.synchronize()
        ms32 = e0.elapsed_time(e1) / n
        fl = 2.0 * npix * M * N
        print(f"{spec}: max-norm rel err {err:.3e} (cuBLAS fp32: {err32:.3e}), elements off by >1%: {bad} of {ref.numel()}; "
              f"{ms * 1e3:.1f} us = {fl / ms / 1e9:.1f} TFLOP/s (cuBLAS fp32 {ms32 * 1e3:.1f} us)")
        if bad:
            d = (dW.double() - ref).abs() > 1e-2 * ref.abs().max()
            print("   bad rows (m) count per 32:", d.sum(1).view(-1, 32).sum(1).tolist()[:16])
            print("   bad cols (n) count per 16:", d.sum(0).view(-1, 16).sum(1).tolist()[:16])
            print("   sample ours/ref:", dW[0, :6].tolist(), ref[0, :6].tolist())
        return err < (2e-3 if npass == 1 else 2e-6)
    B, C, H, W, width, npass = a
    from fincflow_b200.flows import Coupling

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    Coupling.tensor_core = False   # `cp(x)` below is the PyTorch / cuDNN baseline; our path is called through _native
    cp = Coupling((C, H, W), width=width).to(dev)
    with torch.no_grad():  # the last conv is zero-initialised: give it something to do
        cp.net[4].weight.normal_(0, 0.02)
        cp.net[4].bias.normal_(0, 0.1)
        cp.net[4].logs.normal_(0, 0.1)
    x = torch.randn(B, C, H, W, device=dev)
    net = cp.net
    blob = _native.coupling_prepare(net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias,
                                    net[4].logs)
    flags = _native.FLAG_TF32_1PASS if npass == 1 else 0
    if kind == "coupling":
        y, ld = _native.coupling_apply(x, blob, width, flags=flags)
        xr, _ = _native.coupling_apply(y, blob, width, reverse=True, flags=flags)
        torch.cuda.synchronize()
        cpd = Coupling((C, H, W), width=width).to(dev).double()
        cpd.load_state_dict({k: v.double() for k, v in cp.state_dict().items()})
        with torch.no_grad():
            yr, ldr = cpd(x.double())
        ey = float((y.double() - yr).abs().max() / yr.abs().max())
        el = float((ld.double() - ldr).abs().max() / ldr.abs().max().clamp_min(1e-30))
        rt = float((xr - x).abs().max())
        with torch.no_grad():
            y32, ld32 = cp(x)
        print(f"   (PyTorch fp32 vs fp64: y {float((y32.double() - yr).abs().max() / yr.abs().max()):.3e}, "
              f"logdet {float((ld32.double() - ldr).abs().max() / ldr.abs().max()):.3e}; ours vs PyTorch fp32: "
              f"y {float((y - y32).abs().max() / yr.abs().max()):.3e}, logdet {float((ld - ld32).abs().max() / ldr.abs().max()):.3e})")
        print(f"{spec}: y rel {ey:.3e}, logdet rel {el:.3e}, round trip max-abs {rt:.3e}")
        tol = 2e-3 if npass == 1 else 1e-5
        return ey < tol and el < tol and rt < (1e-2 if npass == 1 else 1e-4)
    # timing
    ws = torch.empty(_native.coupling_workspace_bytes(B, C, H, W, width), dtype=torch.uint8, device=dev)
    y = torch.empty_like(x)
    ld = torch.empty(B, device=dev)
    for _ in range(3):
        _native.coupling_apply(x, blob, width, flags=flags, out=y, logdet_out=ld, workspace=ws)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        _native.coupling_apply(x, blob, width, flags=flags, out=y, logdet_out=ld, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flop = 2.0 * B * H * W * (9 * (C // 2) * width + width * width + 9 * width * C)
    with torch.no_grad():
        for _ in range(3):
            cp(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            cp(x)
        e1.record()
        torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / n
    print(f"{spec}: ours {ms:.3f} ms = {flop / ms / 1e9:.1f} TFLOP/s (fp32-equivalent); "
          f"PyTorch/cuDNN fp32 (TF32 off) {ms_t:.3f} ms = {flop / ms_t / 1e9:.1f} TFLOP/s")
    return True


def main():
    if "--case" in sys.argv:
        ok = run_case(sys.argv[sys.argv.index("--case") + 1])
        sys.exit(0 if ok else 1)
    cases = CASES
    if "--only" in sys.argv:
        pref = sys.argv[sys.argv.index("--only") + 1]
        cases = [c for c in CASES if c.startswith(pref)]
    n_fail = 0
    for c in cases:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True,
                               timeout=180)
            out = (p.stdout + p.stderr[-1500:] if p.returncode else p.stdout).strip()
            status = "ok" if p.returncode == 0 else f"FAIL rc={p.returncode}"
        except subprocess.TimeoutExpired:
            out, status = "", "TIMEOUT"
        n_fail += status != "ok"
        print(f"[{status}] ({time.time() - t0:.1f}s) {out}", flush=True)
    print(f"tc_check: {len(cases) - n_fail} / {len(cases)} ok")
    sys.exit(1 if n_fail else 0)


if __name__ == "__main__":
    main()
