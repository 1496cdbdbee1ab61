import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import flows
from fincflow_b200.train import FlowTrainer

def run(flat, steps):
    torch.manual_seed(5)
    m = flows.FastFlow(n_blocks=2, block_size=2, image_size=(3, 16, 16), actnorm=True, width=128).cuda()
    tr = FlowTrainer(m, lr=1e-3, flat_adam=flat)
    m.preprocess.layers[0].fixed_noise = torch.full((8, 3, 16, 16), 0.5, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(9)
    xs = [torch.randint(0, 256, (8, 3, 16, 16), device="cuda", generator=g).float() for _ in range(6)]
    out = []
    for i in range(steps):
        l = float(tr.step(xs[i]))
        out.append((l, {n: p.detach().clone() for n, p in m.named_parameters()}, {n: p.grad.detach().clone() for n, p in m.named_parameters()}))
    return out
a = run(False, 2); b = run(True, 2)
flows.Coupling.tensor_core = False
c = run(False, 2)
flows.Coupling.tensor_core = True
n = 'fastflow_step.0.fastflow_step.glow_unit.glow_step.coupling.net.4.logs'
for name, r in (("torch adam", a), ("flat adam", b), ("pytorch coupling + torch adam", c)):
    print(name, "step-1 logs grad", r[1][2][n][:4].tolist(), "logs", r[1][1][n][:4].tolist(), "loss", r[1][0])
n2 = 'fastflow_step.0.fastflow_step.glow_unit.glow_step.coupling.net.4.weight'
for name, r in (("torch adam", a), ("flat adam", b), ("pytorch coupling + torch adam", c)):
    print(name, "step-0 w", r[0][1][n2].flatten()[:4].tolist(), "step-1 w", r[1][1][n2].flatten()[:4].tolist(), "step-1 gw", r[1][2][n2].flatten()[:4].tolist())
for i in range(2):
    print("step", i, "loss", a[i][0], b[i][0])
    worst = sorted(((float((a[i][1][n] - b[i][1][n]).abs().max() / (a[i][1][n].abs().max() + 1e-12)), n) for n in a[i][1]), reverse=True)[:6]
    print("  params:", worst)
    worstg = sorted(((float((a[i][2][n] - b[i][2][n]).abs().max() / (a[i][2][n].abs().max() + 1e-12)), n) for n in a[i][2]), reverse=True)[:4]
    print("  grads:", worstg)
