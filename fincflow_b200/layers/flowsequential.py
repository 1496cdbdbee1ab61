"""FlowSequential -- the reference's flow container, over the B200 layers.

Operator API of fastflow/layers/flowsequential.py:9-138: `FlowSequential(base_distribution, *modules)`,
`forward(input, context=None, compute_expensive=False) -> (output, logprob + logdet)`,
`log_prob`, `cheap_unnormed_log_prob`, `sample(n) -> (x, x_true)`, `reconstruct(x)`,
`preprocessing_modules()` / `non_preprocessing_modules()` / `non_preprocessing_logdet()`; children are
registered under "0", "1", ... so the reference's state-dict keys (`3.conv_tl.conv.weight`) load.

Semantics kept: layer log-determinants may be python floats (the FInC layers return `0.0`,
fastflow/fastflow.py:34-50) or `[B]` tensors and are summed in layer order; `ModifiedGradFlowLayer`
children receive `compute_expensive`; the base log-prob is added last (flowsequential.py:21-44).

B200 path: when the base distribution is a standard normal (`GaussianPrior`, or any object with
`is_standard_normal = True`) and the output is a CUDA tensor, `logprob + logdet` is ONE launch of
finc_gaussian_logp_f32 (closed form + the running log-determinant) instead of a dense
MultivariateNormal (train/losses.py:17-45) followed by an add; under autograd the same kernel also
produces d logp / d z.  FInC layers built with `logdet_mode="tensor"` add their fused `[B]`
log-determinant like any other tensor-valued layer.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native
from .flowlayer import ModifiedGradFlowLayer, PreprocessingFlowLayer


class _GaussianLogpFn(torch.autograd.Function):
    """logp[n] = -0.5 |z_n|^2 - D/2 log(2 pi) + logdet[n]   (finc_gaussian_logp_f32)"""

    @staticmethod
    def forward(ctx, z, logdet):
        z = z.contiguous()
        ctx.save_for_backward(z)
        ctx.has_logdet = logdet is not None
        logp, _ = _native.gaussian_logp(z, None if logdet is None else logdet.contiguous())
        return logp

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        dz = -z * g.view(-1, *([1] * (z.dim() - 1)))
        return dz, (g if ctx.has_logdet else None)


def _as_batch_tensor(logdet, ref):
    """python floats / 0-dim tensors -> [B] tensor on ref's device (what `+=` broadcasting yields)"""
    if torch.is_tensor(logdet):
        return logdet.expand(ref.shape[0]) if logdet.dim() == 0 else logdet
    return None if logdet == 0 else ref.new_full((ref.shape[0],), float(logdet))


class FlowSequential(nn.Module):
    def __init__(self, base_distribution, *modules):
        super().__init__()
        self.base_distribution = base_distribution
        for name, module in enumerate(modules):
            self.add_module(str(name), module)
        self.sequence_modules = modules

    def __iter__(self):
        return iter(self.sequence_modules)

    def __len__(self):
        return len(self.sequence_modules)

    # ---- one layer, with the reference's dispatch on ModifiedGradFlowLayer --------------------
    @staticmethod
    def _forward_one(module, x, context, compute_expensive):
        if isinstance(module, ModifiedGradFlowLayer):
            return module(x, context, compute_expensive=compute_expensive)
        return module(x, context)

    @staticmethod
    def _reverse_one(module, x, context, compute_expensive):
        if isinstance(module, ModifiedGradFlowLayer):
            out = module.reverse(x, context, compute_expensive)
        else:
            out = module.reverse(x, context)
        # PaddedConv2d.reverse returns (x, 0) in the reference too (layers/conv.py:163); containers
        # written against it index the tensor out, so do we
        return out[0] if isinstance(out, tuple) else out

    def _push(self, modules, x, context, compute_expensive):
        logdet = 0
        for module in modules:
            x, layer_logdet = self._forward_one(module, x, context, compute_expensive)
            logdet = logdet + layer_logdet
        return x, logdet

    def _base_logp(self, z, logdet):
        base = self.base_distribution
        if z.is_cuda and z.dtype == torch.float32 and getattr(base, "is_standard_normal", False):
            return _GaussianLogpFn.apply(z, _as_batch_tensor(logdet, z))
        return base.log_prob(z) + logdet

    # ---- reference API -----------------------------------------------------------------------------
    def forward(self, input, context=None, compute_expensive=False):
        output, logdet = self._push(self.sequence_modules, input, context, compute_expensive)
        return output, self._base_logp(output, logdet)

    def log_prob(self, input, context=None, compute_expensive=True):
        return self.forward(input, context, compute_expensive)[1]

    def cheap_unnormed_log_prob(self, input, context=None):
        return self.log_prob(input, context=context, compute_expensive=False)

    def preprocessing_modules(self):
        return (m for m in self.sequence_modules if isinstance(m, PreprocessingFlowLayer))

    def non_preprocessing_modules(self):
        return (m for m in self.sequence_modules if not isinstance(m, PreprocessingFlowLayer))

    def non_preprocessing_logdet(self, input, context=None, *, compute_expensive=False):
        output, logdet = self._push(list(self.non_preprocessing_modules()), input, context, compute_expensive)
        return self._base_logp(output, logdet)

    def _pull(self, z, context, compute_expensive):
        for module in reversed(self.sequence_modules):
            z = self._reverse_one(module, z, context, compute_expensive)
        return z

    def sample(self, n_samples, context=None, compute_expensive=False, also_true_inverse=False):
        z, _ = self.base_distribution.sample(n_samples, context)
        x = self._pull(z, context, compute_expensive)
        x_true = self._pull(z, context, True) if (also_true_inverse and not compute_expensive) else x
        return x, x_true

    def reconstruct(self, x, context=None, compute_expensive=False):
        z, _ = self._push(self.sequence_modules, x, context, compute_expensive)
        return self._pull(z, context, compute_expensive)
