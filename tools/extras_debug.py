"""run the extra bench workloads one by one with full tracebacks"""
import sys, traceback
sys.path.insert(0, ".")
import torch
import bench
from fincflow_b200.stack import FincStack, HotPathRunner

dev = torch.device("cuda:0")
names = sys.argv[1:] or None
for name, spec in bench.extra_workload_specs(1).items():
    if names and name not in names:
        continue
    try:
        stack = FincStack(spec["levels"]).to(dev)
        r = HotPathRunner(stack, spec["batch"], dev, slots=1, dense_inverse=spec.get("dense", False))
        r.prepare()
        print(name, "ok", "dense levels", sorted(r.dense))
    except Exception:
        print(name, "FAILED")
        traceback.print_exc()
print(bench.run_extra_workloads(torch, dev, 1, 0, None, 5))
