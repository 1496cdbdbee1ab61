"""Golden fixture of a WHOLE multi-scale flow, produced by the reference's own model code
(fastflow/fastflow_cifar_multi_gpu.py::FastFlow) on CPU in the authoring container.

    python tests/golden/make_golden_flow.py       # writes tests/golden/flow_golden.npz

Shims beyond tests/golden/make_golden.py: `utils.convbackward` stub (layers/selfnorm.py JIT-builds
a removed ATen API), `wandb`/`torchvision` stubs if absent, `NegativeGaussianLoss` replaced by a
CPU closed-form twin (the reference pins a dense MultivariateNormal to cuda:0,
train/losses.py:25-27), Coupling width default 512 -> 16 to keep the fixture small, and the
dequantisation noise pinned.  The unit inverse inside `reverse` is the reference's
FastFlowUnit.reverse_level1 (Cython solver); there is no GPU here for reverse_level2.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    PaddedConv2d, FastFlowUnit, solver = import_reference()
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
    import layers.coupling as coupling

    coupling.Coupling.__init__.__defaults__ = (16, None)  # width, n_context
    import train.losses as losses
    import math

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
    try:
        import fastflow_cifar_multi_gpu as ref
    except Exception as e:  # datasets / experiment imports may need more stubs
        raise SystemExit(f"cannot import the reference model script: {e!r}")
    ref.NegativeGaussianLoss = CpuGaussian
    ref.SplitPrior.__init__.__defaults__ = (16,)  # width of the split prior's coupling, 512 -> 16
    ref.FastFlowUnit.reverse = lambda self, x, context=None: self.reverse_level1(x)

    torch.manual_seed(7)
    np.random.seed(7)
    model = ref.FastFlow(n_blocks=3, block_size=2, image_size=(3, 16, 16), actnorm=True)
    B = 4
    x = torch.randint(0, 256, (B, 3, 16, 16)).float()
    noise = torch.rand(B, 3, 16, 16)
    model.preprocess.layers[0].distribution.sample = lambda n, ctx=None: (noise, torch.zeros(n))
    zs, logp = model(x)                                   # first call initialises ActNorm
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.zero_grad()
    # break the zero-init of the coupling output convs so that the flow is non-trivial
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.endswith("net.4.weight"):
                p_.normal_(0, 0.02)
            if n_.endswith("net.4.bias"):
                p_.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    zs, logp = model(x)
    (-logp.sum() / B).backward()
    model.apply(ref.clear_grad)
    grads = {n_: p_.grad.clone() for n_, p_ in model.named_parameters() if "fastflow_unit" in n_}
    with torch.no_grad():
        x_rec = model.reverse(n_samples=B, zs=[z.detach() for z in zs])
    out = {"x": x.numpy(), "noise": noise.numpy(), "logp": logp.detach().numpy(), "x_rec": x_rec.numpy()}
    for i, z in enumerate(zs):
        out[f"zs/{i}"] = z.detach().numpy()
    for k, v in sd.items():
        out[f"sd/{k}"] = v.numpy()
    for k, v in grads.items():
        out[f"grad/{k}"] = v.numpy()
    path = os.path.join(HERE, "flow_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; logp {logp.detach().numpy()}; "
          f"reconstruction max-abs {float((x_rec - x).abs().max())}; {len(sd)} state-dict entries")
    print("\n".join(list(sd.keys())[:12]))


if __name__ == "__main__":
    main()
