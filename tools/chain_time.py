"""chain kernel vs per-unit launches per level shape (run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

dev = torch.device("cuda:0")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for C, H, W in ((12, 16, 16), (24, 8, 8), (48, 4, 4)):
    for U in (1, 16):
        w = torch.stack([FastFlowUnit(C, C, (3, 3)).weight.detach() for _ in range(U)]).to(dev).contiguous()
        x = torch.randn(B, C, H, W, device=dev)
        out = torch.empty(U, B, C, H, W, device=dev)
        bufs = [torch.empty_like(x) for _ in range(U)]
        def per_unit():
            cur = x
            for u in range(U):
                _native.forward(cur, w[u], want_logdet=False, out=bufs[u]); cur = bufs[u]
        t_c = timed(lambda: _native.chain(x, w, out))
        t_u = timed(per_unit)
        A = torch.eye(C, device=dev).repeat(U, 1, 1).contiguous(); b = torch.zeros(U, C, device=dev)
        t_a = timed(lambda: _native.chain(x, w, out, A=A, bias=b))
        def per_unit_aff():
            cur = x
            for u in range(U):
                _native.forward(cur, w[u], want_logdet=False, out=bufs[u]); cur = _native.affine1x1(bufs[u], A[u], b[u])
        t_ua = timed(per_unit_aff)
        print(f"B={B} [{C},{H},{W}] U={U}: chain {t_c:7.1f} us, per-unit {t_u:7.1f} us | with affine: chain {t_a:7.1f} us, FInC + affine1x1 launches {t_ua:7.1f} us")
