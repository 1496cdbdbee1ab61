import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit
dev = torch.device("cuda:0")
C, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (12, 16, 16)
B, U = 256, 16
w = torch.stack([FastFlowUnit(C, C, (3, 3)).weight.detach() for _ in range(U)]).to(dev).contiguous()
x = torch.randn(B, C, H, W, device=dev)
out = torch.empty(U, B, C, H, W, device=dev)
for _ in range(3):
    _native.chain(x, w, out)
torch.cuda.synchronize()
