import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests", "golden")); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from param_fill import filled_state_dict
from fincflow_b200 import flows
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = np.load(os.path.join(REPO, "tests", "golden", "flow_full_golden.npz"))
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for tag, seed, mk in (("cfg2", 21, flows.fastflow_mnist), ("cfg3", 22, flows.fastflow_cifar10)):
    for tc in (False, True):
        for fused_pre in (False, True):
            flows.Coupling.tensor_core = tc
            flows.Preprocess.fused = fused_pre
            model = mk(actnorm=True)
            res = model.load_state_dict(filled_state_dict(model, seed), strict=False)
            model = model.cuda().eval()
            x = torch.from_numpy(g[f"{tag}/x"]).cuda()
            model.preprocess.layers[0].fixed_noise = torch.from_numpy(g[f"{tag}/noise"]).cuda()
            with torch.no_grad():
                zs, logp = model(x)
            errs = [rel(z.cpu().numpy(), g[f"{tag}/zs/{i}"]) for i, z in enumerate(zs)]
            print(tag, "tensor_core", tc, "fused_pre", fused_pre, "missing", res.missing_keys[:3], "unexpected", res.unexpected_keys[:3],
                  "zs rel", ["%.2e" % e for e in errs], "logp rel %.2e" % rel(logp.cpu().numpy(), g[f"{tag}/logp"]))
import pytest
sys.exit(pytest.main(["-x", "-q", os.path.join(REPO, "tests", "test_flowsequential.py"), "-m", "gpu", "--tb=short"]))
