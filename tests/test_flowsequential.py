"""FlowSequential (fincflow_b200/layers/flowsequential.py) against the semantics of the reference's
container (fastflow/layers/flowsequential.py:21-47,89-138): host logic on CPU with toy layers,
the fused base-log-prob path and a reference-style FInC stack on the GPU."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err
from fincflow_b200 import FlowLayer, FlowSequential, ModifiedGradFlowLayer, PreprocessingFlowLayer


class _Scale(FlowLayer):
    """y = a x; logdet as a [B] tensor"""

    def __init__(self, a):
        super().__init__()
        self.a = a

    def forward(self, input, context=None):
        return input * self.a, self.logdet(input)

    def reverse(self, input, context=None):
        return input / self.a

    def logdet(self, input, context=None):
        return torch.full((input.shape[0],), input[0].numel() * math.log(abs(self.a)))


class _Shift(FlowLayer):
    """y = x + b; logdet as the python float 0.0, like the FInC layers (fastflow/fastflow.py:34-50)"""

    def __init__(self, b):
        super().__init__()
        self.b = b

    def forward(self, input, context=None):
        return input + self.b, 0.0

    def reverse(self, input, context=None):
        return input - self.b

    def logdet(self, input, context=None):
        return 0.0


class _Pre(_Shift, PreprocessingFlowLayer):
    pass


class _Modified(ModifiedGradFlowLayer):
    def __init__(self):
        super().__init__()
        self.seen = []

    def forward(self, input, context=None, compute_expensive=False):
        self.seen.append(("f", compute_expensive))
        return input * 2.0, torch.full((input.shape[0],), input[0].numel() * math.log(2.0))

    def reverse(self, input, context=None, compute_expensive=False):
        self.seen.append(("r", compute_expensive))
        return input / 2.0

    def logdet(self, input, context=None, compute_expensive=False):
        return self.forward(input, context, compute_expensive)[1]


class _Normal:
    def __init__(self, shape):
        self.shape = shape

    def log_prob(self, z, context=None):
        return -0.5 * z.flatten(1).pow(2).sum(1) - 0.5 * z[0].numel() * math.log(2 * math.pi)

    def sample(self, n, context=None):
        torch.manual_seed(3)
        z = torch.randn(n, *self.shape)
        return z, self.log_prob(z)


def test_forward_accumulates_float_and_tensor_logdets_in_layer_order():
    mod = _Modified()
    flow = FlowSequential(_Normal((2, 3)), _Pre(0.5), _Scale(3.0), _Shift(-1.0), mod)
    assert [n for n, _ in flow.named_children()] == ["0", "1", "2", "3"]   # reference state-dict keys
    assert len(list(flow)) == 4
    x = torch.randn(5, 2, 3)
    z, logp = flow(x)
    want_z = ((x + 0.5) * 3.0 - 1.0) * 2.0
    assert torch.allclose(z, want_z)
    want = _Normal((2, 3)).log_prob(want_z) + 6 * math.log(3.0) + 6 * math.log(2.0)
    assert torch.allclose(logp, want, atol=1e-5)
    assert mod.seen == [("f", False)]
    assert torch.allclose(flow.log_prob(x), want, atol=1e-5) and mod.seen[-1] == ("f", True)
    flow.cheap_unnormed_log_prob(x)
    assert mod.seen[-1] == ("f", False)
    assert [type(m) for m in flow.preprocessing_modules()] == [_Pre]
    assert len(list(flow.non_preprocessing_modules())) == 3
    z2 = ((x * 3.0) - 1.0) * 2.0
    assert torch.allclose(flow.non_preprocessing_logdet(x),
                          _Normal((2, 3)).log_prob(z2) + 6 * math.log(6.0), atol=1e-5)


def test_sample_and_reconstruct_run_the_layers_in_reverse():
    mod = _Modified()
    flow = FlowSequential(_Normal((4,)), _Scale(0.5), _Shift(2.0), mod)
    x, x_true = flow.sample(7)
    z, _ = _Normal((4,)).sample(7)
    assert torch.allclose(x, (z / 2.0 - 2.0) / 0.5) and x_true is x
    assert mod.seen == [("r", False)]
    x, x_true = flow.sample(7, also_true_inverse=True)
    assert mod.seen[-2:] == [("r", False), ("r", True)] and torch.allclose(x, x_true)
    xin = torch.randn(3, 4)
    assert torch.allclose(flow.reconstruct(xin), xin, atol=1e-6)


@pytest.mark.gpu
def test_reference_style_finc_stack_on_gpu():
    """FlowSequential([Squeeze, FastFlowUnit, ...]) the way fastflow_mnist.py:46-56 builds it:
    same value as the layer-by-layer evaluation, fused Gaussian base log-prob, exact inverse."""
    from fincflow_b200 import FastFlowUnit, GaussianPrior, PaddedConv2d
    from fincflow_b200.flows import Squeeze

    torch.manual_seed(0)
    layers = [Squeeze(), FastFlowUnit(4, 4, (3, 3)), FastFlowUnit(4, 4, (3, 3), logdet_mode="tensor"),
              Squeeze(), FastFlowUnit(16, 16, (3, 3)), PaddedConv2d(16, 16, (3, 3), order="BR")]
    flow = FlowSequential(GaussianPrior((16, 7, 7)), *layers).cuda()
    x = torch.randn(6, 1, 28, 28, device="cuda", requires_grad=True)
    z, logp = flow(x)
    h, ld = x.detach(), 0
    for layer in layers:
        h, l = layer(h)
        ld = ld + l
    assert torch.equal(z.detach(), h)
    want = -0.5 * h.flatten(1).pow(2).sum(1) - 0.5 * 784 * math.log(2 * math.pi) + ld
    assert rel_err(logp.detach().cpu().numpy(), want.detach().cpu().numpy()) <= 1e-6
    # autograd through the fused log-prob kernel == autograd through the closed form
    (-logp.sum() / 6).backward()
    x2 = x.detach().clone().requires_grad_(True)
    h = x2
    for layer in layers:
        h, _ = layer(h)
    (0.5 * h.flatten(1).pow(2).sum() / 6).backward()
    assert rel_err(x.grad.cpu().numpy(), x2.grad.cpu().numpy()) <= 1e-5
    with torch.no_grad():
        assert float((flow.reconstruct(x) - x).abs().max()) <= 1e-4
        xs, xt = flow.sample(5)
        assert xs.shape == (5, 1, 28, 28) and xt is xs
        zs, _ = flow(xs)
        assert np.isfinite(zs.cpu().numpy()).all()
