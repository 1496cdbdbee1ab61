"""Probe: does running the sampling pass (inverse graph) of step k on a second stream, concurrently with
the forward / backward graphs of step k+1, raise the step throughput?  Uses HotPathRunner's phase graphs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fincflow_b200.stack import FincStack, HotPathRunner, cifar10_levels

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, NS, K = 256, 3, 200
stack = FincStack(cifar10_levels(16, 3)).to(dev)
runner = HotPathRunner(stack, B, dev, slots=NS)
g = torch.Generator(device=dev).manual_seed(1)
for s in runner.slots:
    for li in range(len(stack.levels)):
        s.acts[li][0].normal_(generator=g)
        s.zin[li].normal_(generator=g)
runner.prepare()
main = torch.cuda.current_stream(dev)
samp = torch.cuda.Stream(dev)


def serial(n):
    for i in range(n):
        runner.step(i % NS)


def overlapped(n):
    ev_opt = [torch.cuda.Event() for _ in range(n)]
    ev_samp = [torch.cuda.Event() for _ in range(n)]
    for i in range(n):
        sl = i % NS
        runner.run_phase(sl, 0)
        runner.run_phase(sl, 1)
        if i > 0:
            main.wait_event(ev_samp[i - 1])      # the previous sampling pass has finished reading the tables
        runner.run_phase(sl, 2)
        ev_opt[i].record(main)
        samp.wait_event(ev_opt[i])
        with torch.cuda.stream(samp):
            runner.run_phase(sl, 3)
            ev_samp[i].record(samp)
    main.wait_stream(samp)


for name, fn in (("serial", serial), ("overlapped", overlapped), ("serial", serial), ("overlapped", overlapped)):
    fn(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{name:10s}: {ms:.4f} ms/step  {B / ms * 1e3:.0f} images/s", flush=True)
