"""where does a whole-flow step spend its time?  (run under gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from fincflow_b200 import flows

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 256
m = flows.fastflow_cifar10(actnorm=True).to(dev)
x = torch.randint(0, 256, (B, 3, 32, 32), device=dev).float()
opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)


def train():
    opt.zero_grad(set_to_none=True)
    _, logp = m(x)
    (-(logp.sum() / B)).backward()
    opt.step()


def evaluate():
    with torch.no_grad():
        m(x)


def sample():
    with torch.no_grad():
        m.sample(B)


which = sys.argv[1] if len(sys.argv) > 1 else "eval"
fn = {"train": train, "eval": evaluate, "sample": sample}[which]
for _ in range(3):
    fn()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    fn()
torch.cuda.synchronize()
print(which, "wall ms/iter", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
