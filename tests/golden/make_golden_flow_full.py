"""Whole-flow fixtures at the BASELINE configurations, produced by the reference's own model code on CPU:

  cfg2  fastflow_mnist_multi_gpu.FastFlow(n_blocks=2, block_size=16, image_size=(1,28,28), actnorm=True),
        coupling width 512, batch 8: latents, exact log-likelihood, bits per dimension
  cfg3  fastflow_cifar_multi_gpu.FastFlow(n_blocks=3, block_size=16, image_size=(3,32,32), actnorm=True),
        coupling width 512, batch 4

    python tests/golden/make_golden_flow_full.py      # writes tests/golden/flow_full_golden.npz

Parameters come from tests/golden/param_fill.py (a function of the state-dict key), so only inputs and
outputs are stored.  Shims as in make_golden_flow.py (CPU closed-form Gaussian instead of the cuda:0-pinned
MultivariateNormal, reverse_level1 = the reference's Cython solver, dequantisation noise pinned).
"""
import importlib
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402
from param_fill import filled_state_dict  # noqa: E402


def _shims():
    import_reference()
    for name in ("wandb", "torchvision", "torchvision.utils", "torchvision.datasets", "torchvision.transforms"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
    cb = types.ModuleType("utils.convbackward")
    cb.conv2d_backward = None
    sys.modules["utils.convbackward"] = cb
    import train.losses as losses

    class CpuGaussian(torch.nn.Module):
        def __init__(self, size):
            super().__init__()
            self.size, self.dim = size, int(np.prod(size))

        def log_prob(self, input, context=None, sum=True):
            return -0.5 * input.reshape(-1, self.dim).pow(2).sum(1) - 0.5 * self.dim * math.log(2 * math.pi)

        def sample(self, n, context=None):
            x = torch.randn(n, *self.size)
            return x, self.log_prob(x)

    losses.NegativeGaussianLoss = CpuGaussian
    return CpuGaussian


def run(script, kwargs, B, seed, tag, out):
    Gauss = _shims()
    ref = importlib.import_module(script)
    ref.NegativeGaussianLoss = Gauss
    ref.FastFlowUnit.reverse = lambda self, x, context=None: self.reverse_level1(x)
    torch.manual_seed(seed)
    model = ref.FastFlow(**kwargs)
    res = model.load_state_dict(filled_state_dict(model, seed), strict=False)
    C, H, W = kwargs["image_size"]
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randint(0, 256, (B, C, H, W), generator=g).float()
    noise = torch.rand(B, C, H, W, generator=g)
    model.preprocess.layers[0].distribution.sample = lambda n, ctx=None: (noise, torch.zeros(n))
    model.eval()
    with torch.no_grad():
        zs, logp = model(x)
        x_rec = model.reverse(n_samples=B, zs=[z for z in zs])
    bpd = -logp / (math.log(2) * C * H * W)
    out[f"{tag}/x"], out[f"{tag}/noise"] = x.numpy(), noise.numpy()
    out[f"{tag}/logp"], out[f"{tag}/bpd"], out[f"{tag}/x_rec"] = logp.numpy(), bpd.numpy(), x_rec.numpy()
    for i, z in enumerate(zs):
        out[f"{tag}/zs/{i}"] = z.numpy()
    print(f"{tag}: {sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters, unfilled keys "
          f"{[k for k in res.missing_keys][:6]}, logp {logp.numpy()}, bpd {bpd.numpy()}, "
          f"reconstruction max-abs {float((x_rec - x).abs().max())}")


def main():
    out = {}
    run("fastflow_mnist_multi_gpu", dict(n_blocks=2, block_size=16, image_size=(1, 28, 28), actnorm=True), 8, 21, "cfg2", out)
    run("fastflow_cifar_multi_gpu", dict(n_blocks=3, block_size=16, image_size=(3, 32, 32), actnorm=True), 4, 22, "cfg3", out)
    path = os.path.join(HERE, "flow_full_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
