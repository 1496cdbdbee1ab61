"""diagnostic for the weight-gradient GEMM operand layouts (run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native

dev = "cuda"
print("FINC_WGRAD_DEBUG =", os.environ.get("FINC_WGRAD_DEBUG"))
npix, M, N = 64, 128, 64
f = _native.FLAG_TF32_1PASS
P1, Q1 = torch.ones(npix, M, device=dev), torch.ones(npix, N, device=dev)
lib = _native.load()
ws = torch.full((lib.finc_tc_wgrad_workspace_bytes(npix, M, N) // 4,), float("nan"), device=dev)
dW = torch.full((M, N), float("nan"), device=dev)
rc = lib.finc_tc_wgrad_f32(P1.data_ptr(), Q1.data_ptr(), dW.data_ptr(), ws.data_ptr(), ws.numel() * 4, npix, M, N, M, N, N, f, None)
torch.cuda.synchronize()
print("rc", rc, "workspace NaNs left:", int(torch.isnan(ws).sum()), "of", ws.numel(), "; ws[:4]", ws[:4].tolist(), "dW[0,:4]", dW[0, :4].tolist())
print("ones x ones (expect 64):", _native.tc_wgrad(P1, Q1, flags=f)[:2, :4].tolist())
P = torch.arange(npix * M, device=dev, dtype=torch.float32).reshape(npix, M) % 7
print("P x ones: ours", _native.tc_wgrad(P, Q1, flags=f)[:4, 0].tolist(), "ref", (P.t() @ Q1)[:4, 0].tolist())
Q = torch.arange(npix * N, device=dev, dtype=torch.float32).reshape(npix, N) % 5
d = _native.tc_wgrad(P1, Q, flags=f)
print("ones x Q: ours", d[0, :8].tolist(), "ref", (P1.t() @ Q)[0, :8].tolist())
# which Q element feeds output column n?  one-hot Q
for (p, n) in ((0, 0), (1, 0), (8, 0), (0, 1), (0, 4), (0, 32), (3, 37)):
    Qh = torch.zeros(npix, N, device=dev)
    Qh[p, n] = 1.0
    d = _native.tc_wgrad(P1, Qh, flags=f)
    nz = d[0].nonzero().flatten().tolist()
    print(f"one-hot Q[{p},{n}] -> nonzero output columns {nz} values {d[0, nz].tolist()}")
for (p, m) in ((0, 0), (5, 0), (0, 3), (9, 100)):
    Ph = torch.zeros(npix, M, device=dev)
    Ph[p, m] = 1.0
    d = _native.tc_wgrad(Ph, Q1, flags=f)
    nz = d[:, 0].nonzero().flatten().tolist()
    print(f"one-hot P[{p},{m}] -> nonzero output rows {nz}")
