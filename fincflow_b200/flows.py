"""Multi-scale FInC flows (MNIST / CIFAR-10 / ImageNet32 / ImageNet64 shapes).

The model definition of the reference's maintained scripts
(fastflow/fastflow_{mnist,cifar,imagenet,imagenet64}_multi_gpu.py: Preprocess, GlowStep,
FastFlowStep, FastFlowLevel, FastFlow) rebuilt around the B200 FInC kernels:

  * `FastFlowUnit` and `Squeeze` run on the sm_100a kernels behind the C ABI;
  * the Glow glue (ActNorm, Conv1x1, Coupling, SplitPrior, preprocessing) stays plain PyTorch
    ("next" rows) with the reference's formulas
    (layers/actnorm.py:14-64, conv1x1.py:9-43, coupling.py:9-105, normalize.py:18-31,
    transforms.py:11-18, dequantize.py:13-19);
  * the base density is the closed form -0.5|z|^2 - d/2 log(2 pi) on the input's own device
    instead of a dense d x d MultivariateNormal pinned to cuda:0 (train/losses.py:17-45).

Sub-module and parameter names are the reference's, so its checkpoints load with
`load_state_dict` (e.g. `fastflow_levels.0.fastflow_level.1.fastflow_step.fastflow_unit.conv_tl.conv.weight`,
`...glow_unit.glow_step.conv1x1.W`, `...coupling.net.4.logs`).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .fastflow import FastFlowUnit
from .layers.flowlayer import FlowLayer, PreprocessingFlowLayer


# ---------------------------------------------------------------------------------------------
# glue layers
# ---------------------------------------------------------------------------------------------
class _SqueezeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _native.squeeze(x)

    @staticmethod
    def backward(ctx, g):
        return _native.unsqueeze(g)


class _UnsqueezeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _native.unsqueeze(x)

    @staticmethod
    def backward(ctx, g):
        return _native.squeeze(g)


class Squeeze(FlowLayer):
    """space-to-depth, out channel 4c + 2dh + dw (layers/squeeze.py:5-39) on finc_squeeze_f32"""

    def forward(self, input, context=None):
        return _SqueezeFn.apply(input), self.logdet(input, context)

    def reverse(self, input, context=None):
        return _UnsqueezeFn.apply(input)

    def logdet(self, input, context=None):
        return input.new_zeros(len(input))


class Dequantization(PreprocessingFlowLayer):
    """x + u, u ~ U[0,1) (layers/dequantize.py:13-19).  `fixed_noise` pins u for parity tests
    (the reference draws it from the CPU generator on every forward)."""

    def __init__(self):
        super().__init__()
        self.fixed_noise = None

    def forward(self, input, context=None):
        u = self.fixed_noise if self.fixed_noise is not None else torch.rand_like(input)
        return input + u.to(input.device), input.new_zeros(len(input))

    def reverse(self, input, context=None):
        return input.floor()

    def logdet(self, input, context=None):
        return input.new_zeros(len(input))


class Normalization(PreprocessingFlowLayer):
    def __init__(self, translation, scale):
        super().__init__()
        self.register_buffer("translation", torch.Tensor([translation]))
        self.register_buffer("scale", torch.Tensor([scale]))

    def forward(self, input, context=None):
        return (input - self.translation) / self.scale, self.logdet(input, context)

    def reverse(self, input, context=None):
        return input * self.scale + self.translation

    def logdet(self, input, context=None):
        N, C, H, W = input.size()
        return (-C * H * W * torch.log(self.scale)).expand(N)


class LogitTransform(PreprocessingFlowLayer):
    def forward(self, input, context=None):
        return torch.log(input) - torch.log(1 - input), self.logdet(input, context)

    def reverse(self, input, context=None):
        return torch.sigmoid(input)

    def logdet(self, input, context=None):
        return (-torch.log(input) - torch.log(1 - input)).flatten(start_dim=1).sum(-1)


class ActNorm(FlowLayer):
    """layers/actnorm.py:14-64.  `sync_init`: under torch.distributed the data-dependent initialisation uses
    the moments of the GLOBAL batch (all-reduced count / sum / sum of squares), so that every replica of a
    data-parallel job starts with the same translation / log_scale (the reference initialises from
    replica 0's shard inside nn.DataParallel: layers/actnorm.py:17-23)."""

    sync_init = True

    def __init__(self, n_dims):
        super().__init__()
        self.n_dims = n_dims
        self.translation = nn.Parameter(torch.zeros(n_dims))
        self.log_scale = nn.Parameter(torch.zeros(n_dims))
        self.register_buffer("initialized", torch.tensor(0))
        self._ready = False   # host copy of `initialized`: checking the device buffer costs a stream sync per call

    def is_initialized(self):
        if not self._ready:
            self._ready = bool(self.initialized)
        return self._ready

    @staticmethod
    def _moments(input):
        dims = [0, 2, 3]
        import torch.distributed as dist

        if ActNorm.sync_init and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            n = torch.tensor([float(input.numel() // input.shape[1])], device=input.device, dtype=torch.float64)
            s = input.double().sum(dim=dims)
            ss = input.double().pow(2).sum(dim=dims)
            packed = torch.cat([n, s, ss])
            dist.all_reduce(packed)
            n, s, ss = packed[0], packed[1:1 + s.numel()], packed[1 + s.numel():]
            mean = s / n
            var = (ss - n * mean * mean) / (n - 1)          # unbiased, like torch.std
            return mean.to(input.dtype), var.clamp_min(0).sqrt().to(input.dtype)
        return input.mean(dim=dims), input.std(dim=dims)

    def forward(self, input, context=None):
        if not self.is_initialized():  # data-dependent init on the first batch (actnorm.py:17-23)
            with torch.no_grad():
                mean, std = self._moments(input)
                self.translation.data.copy_(mean)
                self.log_scale.data.copy_(torch.log(std + 1e-8))
                self.initialized.fill_(1)
                self._ready = True
        t, ls = self.translation.view(1, -1, 1, 1), self.log_scale.view(1, -1, 1, 1)
        return (input - t) * torch.exp(-ls), self.logdet(input, context)

    def reverse(self, input, context=None):
        assert self.is_initialized()
        t, ls = self.translation.view(1, -1, 1, 1), self.log_scale.view(1, -1, 1, 1)
        return input * torch.exp(ls) + t

    def logdet(self, input, context=None):
        H, W = input.shape[2:]
        return -self.log_scale.sum().expand(input.size(0)) * H * W


class Conv1x1(FlowLayer):
    def __init__(self, n_channels):
        super().__init__()
        self.n_channels = n_channels
        q = torch.linalg.qr(torch.randn(n_channels, n_channels))[0]
        self.W = nn.Parameter(q.contiguous())

    def forward(self, x, context=None):
        _, _, H, W = x.size()
        ldj = H * W * torch.slogdet(self.W)[1]
        return F.conv2d(x, self.W.view(self.n_channels, self.n_channels, 1, 1)), ldj

    def reverse(self, z, context=None):
        return F.conv2d(z, torch.inverse(self.W).view(self.n_channels, self.n_channels, 1, 1))

    def logdet(self, input, context=None):
        return input.shape[2] * input.shape[3] * torch.slogdet(self.W)[1]


class Conv2dZero(nn.Module):
    def __init__(self, in_channels, out_channels, logscale_factor=3):
        super().__init__()
        self.logscale_factor = logscale_factor
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.logs = nn.Parameter(torch.zeros(out_channels))

    def forward(self, input):
        out = F.conv2d(input, self.weight, self.bias, padding=1)
        return out * torch.exp(self.logs * self.logscale_factor).view(1, -1, 1, 1)


class Coupling(FlowLayer):
    """Affine coupling (layers/coupling.py:44-105).  On CUDA the whole layer -- both 3x3 convolutions,
    the 1x1, the ReLUs, exp(3 logs), the tanh-bounded scale, the affine map and the per-image
    log-determinant -- runs on the tensor-core path (finc_coupling_apply_f32: tcgen05 3xTF32
    implicit GEMMs, hidden activations channels-last, never seen by PyTorch).  Shapes the path does
    not cover (e.g. width % 32 != 0) and CPU tensors use the PyTorch formulas below.

    `precision`: "fp32" = 3xTF32 split products with fp32 accumulation (matches an fp32 reference
    to ~1e-6); "tf32" = single-pass TF32 products, PyTorch's default conv precision (~5e-4)."""

    tensor_core = True      # class-wide switch (tests compare both paths)
    precision = "fp32"

    def __init__(self, input_size, width=512):
        super().__init__()
        self.n_channels = input_size[0]
        self.half_channels = self.n_channels // 2
        self.width = width
        self.net = nn.Sequential(nn.Conv2d(self.half_channels, width, 3, padding=1), nn.ReLU(),
                                 nn.Conv2d(width, width, 1), nn.ReLU(), Conv2dZero(width, self.n_channels))
        self._blob = None
        self._blob_key = None

    # ---- tensor-core path -------------------------------------------------------------------------
    def _params(self):
        n = self.net
        return (n[0].weight, n[0].bias, n[2].weight, n[2].bias, n[4].weight, n[4].bias, n[4].logs)

    def _use_tc(self, x):
        return (self.tensor_core and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                and _native.coupling_prepared_bytes(self.n_channels, self.width) > 0)

    def prepared(self, with_backward=False, training=False):
        """hi / lo split weight blob.  Evaluation (no_grad) reuses it while no parameter changed: the cache is
        keyed on (data_ptr, `_version`) of every parameter -- in-place updates, load_state_dict and most optimizers
        bump `_version`.  torch.optim.Adam(fused=True) does NOT (measured: version 0 -> 0 after step()), so a
        training forward (`training=True`) never trusts the cache: it always rebuilds the blob and leaves the
        cache invalid, and the first evaluation after an optimizer step rebuilds it once more.  `with_backward`
        adds the transposed weights of the backward pass."""
        ps = self._params()
        key = None if training else (with_backward,) + tuple((p.data_ptr(), p._version) for p in ps)
        if training or self._blob is None or key != self._blob_key or self._blob.device != ps[0].device:
            self._blob = _native.coupling_prepare(*[p.detach() for p in ps], self.net[4].logscale_factor,
                                                  out=self._blob if self._blob is not None
                                                  and self._blob.device == ps[0].device else None,
                                                  with_backward=with_backward)
            self._blob_key = key
        return self._blob

    def _flags(self):
        return _native.FLAG_TF32_1PASS if self.precision == "tf32" else 0

    # ---- reference formulas (CPU tensors / uncovered shapes / autograd) ----------------------------
    def _st(self, x):
        x1, x2 = x[:, :self.half_channels], x[:, self.half_channels:]
        h = self.net(x1)
        log_s = 2.0 * torch.tanh(h[:, ::2] / 2.0)
        return x1, x2, log_s, h[:, 1::2]

    def forward(self, input, context=None):
        if self._use_tc(input):
            return _CouplingFn.apply(input, self, *self._params())
        x1, x2, log_s, t = self._st(input)
        return torch.cat([x1, x2 * torch.exp(log_s) + t], dim=1), log_s.flatten(start_dim=1).sum(-1)

    def reverse(self, input, context=None):
        if self._use_tc(input) and not torch.is_grad_enabled():
            return _native.coupling_apply(input, self.prepared(), self.width, reverse=True, flags=self._flags())[0]
        x1, x2, log_s, t = self._st(input)
        return torch.cat([x1, (x2 - t) * torch.exp(-log_s)], dim=1)

    def logdet(self, input, context=None):
        return self.forward(input, context)[1]


class _CouplingFn(torch.autograd.Function):
    """Coupling.forward and its backward on the tensor-core kernels (finc_coupling_apply_f32 /
    finc_coupling_backward_f32).  The forward keeps its workspace (the channels-last hidden activations)
    for the backward.  Widths the backward kernels do not cover (width % 128 != 0) recompute through the
    PyTorch formulas instead."""

    @staticmethod
    def forward(ctx, x, module, *params):
        ctx.module = module
        B, C, H, W = x.shape
        train = any(ctx.needs_input_grad)
        ctx.native = train and _native.coupling_prepared_bytes(C, module.width, True) > 0
        ws = None
        if ctx.native:
            ws = torch.empty(_native.coupling_workspace_bytes(B, C, H, W, module.width), dtype=torch.uint8, device=x.device)
        blob = module.prepared(with_backward=ctx.native, training=train)
        y, logdet = _native.coupling_apply(x, blob, module.width, flags=module._flags(), workspace=ws)
        ctx.save_for_backward(x)
        ctx.ws, ctx.blob = ws, blob
        return y, logdet

    @staticmethod
    def backward(ctx, dy, dlogdet):
        (x,) = ctx.saved_tensors
        m = ctx.module
        if ctx.native:
            if dy is None:
                dy = torch.zeros_like(x)
            grads = _native.coupling_backward(x, dy.contiguous(), dlogdet, ctx.blob, ctx.ws, m.width, flags=m._flags())
            ctx.ws = None
            return (grads[0], None, *grads[1:])
        ps = m._params()
        with torch.enable_grad():
            xi = x.detach().requires_grad_(ctx.needs_input_grad[0])
            x1, x2, log_s, t = m._st(xi)
            y = torch.cat([x1, x2 * torch.exp(log_s) + t], dim=1)
            ld = log_s.flatten(start_dim=1).sum(-1)
            wanted = [xi] if ctx.needs_input_grad[0] else []
            wanted += [p for p, need in zip(ps, ctx.needs_input_grad[2:]) if need]
            grads = list(torch.autograd.grad([y, ld], wanted, [dy, dlogdet], allow_unused=True))
        dx = grads.pop(0) if ctx.needs_input_grad[0] else None
        dps = [grads.pop(0) if need else None for need in ctx.needs_input_grad[2:]]
        return (dx, None, *dps)


class GaussianPrior(nn.Module):
    """standard normal over `size`; closed form, device follows the input / `device`"""

    is_standard_normal = True  # FlowSequential then fuses log_prob + logdet into finc_gaussian_logp_f32

    def __init__(self, size):
        super().__init__()
        self.size = tuple(size)
        self.dim = int(math.prod(size))
        self.register_buffer("_anchor", torch.zeros(()), persistent=False)

    def log_prob(self, input, context=None):
        return -0.5 * input.reshape(input.shape[0], -1).pow(2).sum(1) - 0.5 * self.dim * math.log(2 * math.pi)

    def forward(self, input, context=None):
        return -self.log_prob(input, context).sum(-1)

    def sample(self, n_samples, context=None):
        x = torch.randn(n_samples, *self.size, device=self._anchor.device)
        return x, self.log_prob(x)


class SplitPrior(FlowLayer):
    """Coupling, then the second half of the channels is modelled by the prior
    (fastflow_cifar_multi_gpu.py:41-75)"""

    def __init__(self, size, width=512):
        super().__init__()
        self.n_channels = size[0]
        self.transform = Coupling(size, width=width)
        self.base = GaussianPrior((self.n_channels // 2, size[1], size[2]))

    def forward(self, input, context=None):
        x, ldj = self.transform(input, context)
        x1, x2 = x[:, :self.n_channels // 2], x[:, self.n_channels // 2:]
        return x1, x2, self.base.log_prob(x2) + ldj

    def reverse(self, input, x2=None, context=None):
        if x2 is None:
            x2 = torch.randn_like(input)
        return self.transform.reverse(torch.cat([input, x2], dim=1), context)

    def logdet(self, input, context=None):
        return self.forward(input, context)[2]


# ---------------------------------------------------------------------------------------------
# flow
# ---------------------------------------------------------------------------------------------
class Preprocess(nn.Module):
    """Dequantization, Normalization(0, 256), Normalization(-alpha, 1/(1-2 alpha)), LogitTransform
    (fastflow_cifar_multi_gpu.py:162-186).  On CUDA the four layers and their log-determinants are ONE
    kernel (finc_preprocess_f32); the sub-modules stay for the reference's state-dict keys
    (`preprocess.layers.1.translation`, ...) and for CPU tensors."""

    fused = True

    def __init__(self, size):
        super().__init__()
        self.alpha = alpha = 1e-6
        self.layers = nn.Sequential(Dequantization(), Normalization(0, 256),
                                    Normalization(-alpha, 1 / (1 - 2 * alpha)), LogitTransform())

    def forward(self, x):
        if self.fused and x.is_cuda and x.dtype == torch.float32 and not x.requires_grad:
            deq = self.layers[0]
            u = deq.fixed_noise.to(x.device) if deq.fixed_noise is not None else torch.rand_like(x)
            return _native.preprocess(x, u, self.alpha)
        logdet = 0
        for layer in self.layers:
            x, ld = layer(x)
            logdet = logdet + ld
        return x, logdet

    def reverse(self, x):
        if self.fused and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
            return _native.preprocess(x, None, self.alpha, reverse=True)[0]
        for layer in reversed(self.layers):
            x = layer.reverse(x)
        return x


class _Chain(nn.Module):
    def _fwd(self, seq, x):
        logdet = 0
        for layer in seq:
            x, ld = layer(x)
            logdet = logdet + ld
        return x, logdet

    def _rev(self, seq, x):
        for layer in reversed(seq):
            x = layer.reverse(x)
        return x


class GlowStep(_Chain):
    def __init__(self, size, actnorm=False, width=512):
        super().__init__()
        d = OrderedDict()
        if actnorm:
            d["actnorm"] = ActNorm(size[0])
        d["conv1x1"] = Conv1x1(size[0])
        d["coupling"] = Coupling(size, width)
        self.glow_step = nn.Sequential(d)

    fused = True  # ActNorm + Conv1x1 as one finc_affine1x1_f32 launch on CUDA tensors

    def forward(self, x):
        if self.fused and x.is_cuda:
            return _fused_glow_forward(self, x)
        return self._fwd(self.glow_step, x)

    def reverse(self, x):
        if self.fused and x.is_cuda:
            return _fused_glow_reverse(self, x)
        return self._rev(self.glow_step, x)


class _Affine1x1Fn(torch.autograd.Function):
    """y = A x + b per pixel on finc_affine1x1_f32; backward-data = the same kernel with A^T;
    dA = sum dy x^T and db = sum dy on finc_affine1x1_backward_weight_f32."""

    @staticmethod
    def forward(ctx, x, A, b):
        x = x.contiguous()
        A = A.contiguous()
        ctx.save_for_backward(x, A)
        ctx.has_bias = b is not None
        return _native.affine1x1(x, A, None if b is None else b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, A = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _native.affine1x1(dy, A.t().contiguous()) if ctx.needs_input_grad[0] else None
        dA = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dA, db = _native.affine1x1_backward_weight(dy, x, want_bias=ctx.has_bias)
        return dx, dA, db


def affine_of(actnorm, conv1x1):
    """(A, b) of ActNorm followed by Conv1x1: W ((x - t) * exp(-log_s)) = A x + b"""
    W = conv1x1.W
    if actnorm is None:
        return W, None
    A = W * torch.exp(-actnorm.log_scale).unsqueeze(0)
    return A, -(A @ actnorm.translation)


class _SlogdetFn(torch.autograd.Function):
    """log|det W_i| of a stack of Conv1x1 weights on finc_slogdet_inverse_f32; backward g_i * W_i^-T"""

    @staticmethod
    def forward(ctx, Ws):
        ld, Winv = _native.slogdet_inverse(Ws)
        ctx.save_for_backward(Winv)
        return ld

    @staticmethod
    def backward(ctx, g):
        (Winv,) = ctx.saved_tensors
        return g.view(-1, 1, 1) * Winv.transpose(1, 2)


def _glue_constants(self):
    """(A, b, scalar logdet per pixel, A^-1, reverse bias) of [ActNorm +] Conv1x1, cached on the parameter
    versions.  Used under no_grad (evaluation / sampling): slogdet, inverse and the small products then run
    once per weight update instead of once per call (torch.slogdet alone is an LU factorisation + a host sync)."""
    gs = self.glow_step
    act = getattr(gs, "actnorm", None)
    conv = gs.conv1x1
    ps = [conv.W] + ([act.translation, act.log_scale] if act is not None else [])
    key = tuple((p.data_ptr(), p._version) for p in ps)
    if getattr(self, "_glue_key", None) != key:
        with torch.no_grad():
            W = conv.W.detach()
            if W.is_cuda and W.shape[0] <= 128:
                ld, Winv = _native.slogdet_inverse(W.unsqueeze(0).contiguous())
                ld, Winv = ld[0], Winv[0]
            else:
                ld, Winv = torch.slogdet(W)[1], torch.inverse(W)
            if act is None:
                A, b, Ainv, binv = W, None, Winv, None
            else:
                A = W * torch.exp(-act.log_scale.detach()).unsqueeze(0)
                b = -(A @ act.translation.detach())
                ld = ld - act.log_scale.detach().sum()
                Ainv = torch.exp(act.log_scale.detach()).unsqueeze(1) * Winv
                binv = act.translation.detach().contiguous()
            self._glue = (A.contiguous(), None if b is None else b.contiguous(), ld, Ainv.contiguous(), binv)
        self._glue_key = key
    return self._glue


def _fused_glow_forward(self, x):
    """GlowStep.forward with [ActNorm +] Conv1x1 as ONE per-pixel affine kernel;
    same values and log-determinants as the layer-by-layer path (actnorm.py:14-64, conv1x1.py:18-43)."""
    gs = self.glow_step
    act = getattr(gs, "actnorm", None)
    conv = gs.conv1x1
    B, _, H, W = x.shape
    if act is not None and not act.is_initialized():
        x, ld0 = act(x)                     # data-dependent init needs the un-fused statistics once
        y = _Affine1x1Fn.apply(x, conv.W, None)
        y, ld = gs.coupling(y)
        return y, ld0 + H * W * torch.slogdet(conv.W)[1] + ld
    if not torch.is_grad_enabled():
        A, b, ld_pix, _, _ = _glue_constants(self)
        y = _native.affine1x1(x.contiguous(), A, b)
        y, ld = gs.coupling(y)
        return y, ld + (H * W) * ld_pix
    self._glue_key = None   # training forward: the cached evaluation constants may not be trusted afterwards
                            # (torch's fused Adam updates parameters without bumping their version counters)
    A, b = affine_of(act, conv)
    y = _Affine1x1Fn.apply(x, A, b)
    ld_w = conv.__dict__.pop("_ld_batched", None)     # set by FastFlow.forward: one batched LU per level
    if ld_w is None:
        ld_w = torch.slogdet(conv.W)[1]
    ld_pix = ld_w if act is None else ld_w - act.log_scale.sum()
    y, ld = gs.coupling(y)
    return y, ld + (H * W) * ld_pix


def _fused_glow_reverse(self, z):
    gs = self.glow_step
    act = getattr(gs, "actnorm", None)
    assert act is None or act.is_initialized()
    x = gs.coupling.reverse(z)
    _, _, _, Ainv, binv = _glue_constants(self)
    return _native.affine1x1(x.contiguous(), Ainv, binv)


class FastFlowStep(_Chain):
    def __init__(self, size, actnorm=False, kernel_size=(3, 3), width=512, mask_in_backward=True):
        super().__init__()
        # mask_in_backward=True applies the FInC gradient mask inside the weight-gradient kernel.  The reference clips
        # the gradient norm over the UNMASKED gradients and masks afterwards (train/experiment.py:243-250): pass
        # False (and call clear_grad / reset_gradients after clipping) when that ordering matters.
        self.fastflow_step = nn.Sequential(OrderedDict([
            ("fastflow_unit", FastFlowUnit(size[0], size[0], kernel_size, mask_in_backward=mask_in_backward)),
            ("glow_unit", GlowStep(size, actnorm, width))]))

    fused = True  # evaluation: FastFlowUnit + [ActNorm +] Conv1x1 as ONE finc_chain_f32 launch

    def forward(self, x):
        unit, glow = self.fastflow_step.fastflow_unit, self.fastflow_step.glow_unit
        act = getattr(glow.glow_step, "actnorm", None)
        if (self.fused and glow.fused and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32
                and (act is None or act.is_initialized()) and x.is_contiguous() and x.data_ptr() % 16 == 0
                and unit.weight.data_ptr() % 16 == 0 and unit.weight.is_contiguous()
                and _native.chain_supported(4, unit.cq, x.shape[2], x.shape[3], unit.kernel_size, True)):
            # z = FInC(x) never leaves shared memory: y = A z + b is written, then the coupling layer
            A, b, ld_pix, _, _ = _glue_constants(glow)
            if b is None:   # no ActNorm: plain Conv1x1
                b = getattr(self, "_zero_bias", None)
                if b is None or b.device != x.device:
                    b = self._zero_bias = torch.zeros(A.shape[0], device=x.device)
            H, W = x.shape[2], x.shape[3]
            y = _native.chain(x, unit.weight.detach().unsqueeze(0), torch.empty_like(x),
                              A=A.unsqueeze(0), bias=b.unsqueeze(0))
            y, ld = glow.glow_step.coupling(y)
            ld = ld + (H * W) * ld_pix
            if unit.logdet_mode == "tensor":   # otherwise the unit reports 0.0 like the reference (fastflow.py:34-50)
                ld = ld + unit.logdet(x)
            return y, ld
        return self._fwd(self.fastflow_step, x)

    def reverse(self, x):
        return self._rev(self.fastflow_step, x)


class FastFlowLevel(nn.Module):
    def __init__(self, size, block_size=16, actnorm=False, kernel_size=(3, 3), width=512, mask_in_backward=True):
        super().__init__()
        size = (size[0] * 4, size[1] // 2, size[2] // 2)
        self.fastflow_level = nn.ModuleList([Squeeze(),
                                             *[FastFlowStep(size, actnorm, kernel_size, width, mask_in_backward)
                                               for _ in range(block_size)],
                                             SplitPrior(size, width)])

    def forward(self, x):
        logdet, z = 0, None
        for layer in self.fastflow_level:
            if isinstance(layer, SplitPrior):
                x, z, ld = layer(x)
            else:
                x, ld = layer(x)
            logdet = logdet + ld
        return x, z, logdet

    def reverse(self, x, z=None):
        for layer in reversed(self.fastflow_level):
            x = layer.reverse(x, z) if isinstance(layer, SplitPrior) else layer.reverse(x)
        return x


class FastFlow(nn.Module):
    """forward(x) -> (zs, logp[B]); reverse(n_samples, zs) -> x; sample(n); reconstruct(x)
    (fastflow_cifar_multi_gpu.py:295-387).  `final_steps=None` = block_size (CIFAR / ImageNet32
    scripts); the MNIST and ImageNet64 scripts use ONE final step (`final_steps=1`)."""

    def __init__(self, n_blocks=2, block_size=16, image_size=(1, 28, 28), actnorm=False, kernel_size=(3, 3),
                 final_steps=None, width=512, mask_in_backward=True):
        super().__init__()
        C_in, H, W = image_size
        self.output_size = (C_in * 2 ** (n_blocks + 1), H // 2 ** n_blocks, W // 2 ** n_blocks)
        self.preprocess = Preprocess(image_size)
        self.fastflow_levels = nn.ModuleList([
            FastFlowLevel((C_in * 2 ** i, H // 2 ** i, W // 2 ** i), block_size, actnorm, kernel_size, width,
                          mask_in_backward)
            for i in range(n_blocks - 1)])
        self.squeeze = Squeeze()
        n_final = block_size if final_steps is None else final_steps
        self.fastflow_step = nn.Sequential(*[FastFlowStep(self.output_size, actnorm, kernel_size, width, mask_in_backward)
                                             for _ in range(n_final)])
        self.base_distribution = GaussianPrior(self.output_size)

    def _batched_slogdets(self):
        """training path: log|det W| of every Conv1x1 in ONE batched LU per channel count (3 calls for the
        CIFAR-10 flow instead of 48; each torch.slogdet is a cuSOLVER factorisation with a host round trip)"""
        groups = {}
        for mod in self.modules():
            if isinstance(mod, Conv1x1):
                groups.setdefault(mod.n_channels, []).append(mod)
        for n_ch, convs in groups.items():
            Ws = torch.stack([c.W for c in convs])
            lds = _SlogdetFn.apply(Ws) if n_ch <= 128 else torch.linalg.slogdet(Ws)[1]
            for i, c in enumerate(convs):
                c._ld_batched = lds[i]

    def forward(self, x, context=None):
        zs = []
        if torch.is_grad_enabled() and x.is_cuda and GlowStep.fused:
            self._batched_slogdets()
        x, logdet = self.preprocess(x)
        for level in self.fastflow_levels:
            x, z, ld = level(x)
            logdet = logdet + ld
            zs.append(z)
        x, ld = self.squeeze(x)
        logdet = logdet + ld
        for step in self.fastflow_step:
            x, ld = step(x)
            logdet = logdet + ld
        zs.append(x)
        return zs, logdet + self.base_distribution.log_prob(x)

    def reverse(self, n_samples=1, zs=None, z_std=1.0):
        if zs is None:
            zs = [self.base_distribution.sample(n_samples)[0]]
        z = zs[-1]
        for step in reversed(self.fastflow_step):
            z = step.reverse(z)
        x = self.squeeze.reverse(z)
        for i, level in enumerate(reversed(self.fastflow_levels)):
            x = level.reverse(x) if len(zs) == 1 else level.reverse(x, zs[-i - 2])
        return self.preprocess.reverse(x)

    def log_prob(self, x, bits_per_pixel=False):
        """bits_per_pixel: per IMAGE dimension (x[0].numel()).  The reference divides by `zs[0].numel()`, a
        batch-sized latent tensor (fastflow_cifar_multi_gpu.py log_prob), so its printed value scales with the batch
        size and is not comparable; its training metric (train/experiment.py:279-295) agrees with this one."""
        zs, logp = self.forward(x)
        return logp / (math.log(2) * x[0].numel()) if bits_per_pixel else logp

    def sample(self, n_samples, context=None):
        x = self.reverse(n_samples=n_samples)
        return x, x

    def reconstruct(self, x, context=None):
        zs, _ = self.forward(x)
        return self.reverse(n_samples=x.shape[0], zs=[zs[-1]])

    def reconstruct_exact(self, x):
        """true inverse: feeds the split-off latents back (test_layers.py:304-348 reconstruct_ff)"""
        zs, _ = self.forward(x)
        return self.reverse(n_samples=x.shape[0], zs=zs)


class InferenceSession:
    """Likelihood evaluation and sampling of a trained flow as CUDA graphs.

    An eager forward of the CIFAR-10 flow issues ~1900 kernel launches (48 steps x FInC unit, affine glue,
    six coupling launches, log-determinant adds); the host needs ~26 ms to enqueue what the GPU executes in
    ~12 ms.  The session captures `model.forward` / `model.reverse` once per batch size into a CUDA graph
    (static input / output buffers, weights read in place) and replays it.  Weights must not change between
    capture and replay; call `reset()` after loading new ones.  The reference samples one layer at a time
    from Python (train/experiment.py:327-337)."""

    def __init__(self, model):
        self.model = model.eval()
        self._eval, self._sample = {}, {}

    def reset(self):
        self._eval.clear()
        self._sample.clear()

    @staticmethod
    def _capture(fn, warm=2):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warm):          # lazy initialisations (weight blobs, glue constants, workspaces)
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            out = fn()
        return g, out

    def log_prob(self, x):
        """(zs, logp[B]) like model.forward(x); the returned tensors are the session's static buffers"""
        key = tuple(x.shape)
        if key not in self._eval:
            buf = x.clone()
            g, out = self._capture(lambda: self.model(buf))
            self._eval[key] = (g, buf, out)
        g, buf, out = self._eval[key]
        buf.copy_(x)
        g.replay()
        return out

    def sample(self, n_samples):
        """x [n, C, H, W] like model.sample(n)[0]; latents are drawn inside the graph"""
        if n_samples not in self._sample:
            g, out = self._capture(lambda: self.model.sample(n_samples)[0])
            self._sample[n_samples] = (g, out)
        g, out = self._sample[n_samples]
        g.replay()
        return out


def set_fp32_parity(enabled: bool = True):
    """The reference ran full fp32 (CUDA 10.2 / cuDNN 7.6 predate TF32).  The glue layers here
    are PyTorch convolutions whose cuDNN default is TF32 (~5e-4 relative); call this to get the
    reference's numerics in the glue as well.  The FInC kernels are always fp32 FMA."""
    torch.backends.cudnn.allow_tf32 = not enabled
    torch.backends.cuda.matmul.allow_tf32 = not enabled


def clear_grad(module):
    from .fastflow import clear_grad as _cg

    _cg(module)


# builders with the reference scripts' settings (SURVEY.md section 8 shape table)
def fastflow_mnist(**kw):
    return FastFlow(n_blocks=2, block_size=16, image_size=(1, 28, 28), final_steps=1, **kw)


def fastflow_cifar10(**kw):
    return FastFlow(n_blocks=3, block_size=16, image_size=(3, 32, 32), **kw)


def fastflow_imagenet32(**kw):
    return FastFlow(n_blocks=3, block_size=48, image_size=(3, 32, 32), **kw)


def fastflow_imagenet64(kernel_size=(3, 3), **kw):
    return FastFlow(n_blocks=4, block_size=48, image_size=(3, 64, 64), final_steps=1, kernel_size=kernel_size, **kw)
