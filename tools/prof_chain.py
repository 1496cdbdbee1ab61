"""The round-2 launches of the headline step, eagerly, at the bench's three level shapes (batch 256, 16 units):
forward chain, backward-data chain, batched dW, in-place inverse chain -- the command profiled under ncu
(profiles/r2_chain_full_b256.md) -- plus one tensor-core Coupling forward / backward at level 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native
from fincflow_b200.fastflow import FastFlowUnit

dev = torch.device("cuda:0")
B, U, reps = 256, 16, (1 if "--once" in sys.argv else 2)
torch.manual_seed(0)
for C, H, W in ((12, 16, 16), (24, 8, 8), (48, 4, 4)):
    w = torch.stack([FastFlowUnit(C, C, (3, 3)).weight.detach() for _ in range(U)]).to(dev).contiguous()
    x = torch.randn(B, C, H, W, device=dev)
    acts = torch.empty(U, B, C, H, W, device=dev)
    dzs = torch.randn(U + 1, B, C, H, W, device=dev)
    dw = torch.empty_like(w)
    nb = _native.prepared_weights_bytes(_native.PREP_INVERSE, B, 4, C // 4, H, W, 3, 3)
    tables = torch.empty((U, nb), dtype=torch.uint8, device=dev)
    _native.prepare_weights(w, tables, _native.PREP_INVERSE, B, H, W)
    ws = torch.zeros(_native.backward_weight_batched_workspace_bytes(B, 4, C // 4, H, W, 3, 3, U - 1), dtype=torch.uint8, device=dev)
    out = torch.empty_like(x)
    for _ in range(reps):
        _native.chain(x, w, acts)
        _native.chain(dzs[U], w, dzs[:U], units=range(U - 1, 0, -1), transpose=True)
        _native.backward_weight_batched(dzs[2:], acts[:U - 1], dw[1:], (3, 3), workspace=ws)
        _native.inverse_chain(x, tables, (3, 3), range(U - 1, -1, -1), out=out)
    torch.cuda.synchronize()
if "--coupling" in sys.argv:
    from fincflow_b200.flows import Coupling
    cp = Coupling((12, 16, 16), width=512).to(dev)
    xc = torch.randn(B, 12, 16, 16, device=dev, requires_grad=True)
    for _ in range(reps):
        y, ld = cp(xc)
        (y.sum() + ld.sum()).backward()
    torch.cuda.synchronize()
print("ok")
