"""torch.autograd glue over the C ABI: one Function for the fused FInC convolution."""
from __future__ import annotations

import torch

from . import _native


class FincConvFunction(torch.autograd.Function):
    """z, logdet = FInC conv of x [B, G*C, H, W] with packed weights [G*C, C, kH, kW].

    forward  -> finc_forward_f32            (reference: layers/conv.py:102-107, fastflow.py:31-50)
    backward -> finc_backward_input_f32 + finc_backward_weight_f32
                (reference: cuDNN dgrad/wgrad via autograd; mask: layers/conv.py:98-99)
    `mask_dw=True` applies the FInC gradient mask inside the weight-gradient kernel;
    `False` returns the raw gradient exactly as autograd does in the reference, to be
    masked later by reset_gradients()/clear_grad (train/experiment.py:246-250 masks AFTER
    the optional gradient clipping).
    """

    @staticmethod
    def forward(ctx, x, w, G, orders, mask_dw, want_logdet):
        z, logdet = _native.forward(x, w, G, orders, want_logdet)
        ctx.save_for_backward(x, w)
        ctx.G, ctx.orders, ctx.mask_dw = G, orders, mask_dw
        if logdet is None:
            logdet = x.new_zeros(())
        ctx.mark_non_differentiable(logdet)
        return z, logdet

    @staticmethod
    def backward(ctx, dz, _dlogdet):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _native.backward_input(dz, w, ctx.G, ctx.orders)
        if ctx.needs_input_grad[1]:
            flags = 0 if ctx.mask_dw else _native.FLAG_NO_MASK
            dw = _native.backward_weight(dz, x, (w.shape[2], w.shape[3]), ctx.G, ctx.orders, flags)
        return dx, dw, None, None, None, None


def finc_conv(x, w, G=4, orders=_native.ORDERS_UNIT, mask_dw=False, want_logdet=True):
    z, logdet = FincConvFunction.apply(x, w, G, orders, mask_dw, want_logdet)
    return z, (logdet if want_logdet else None)


def finc_inverse(z, w, G=4, orders=_native.ORDERS_UNIT):
    """x with forward(x) == z; no autograd (the reference's reverse has none either)."""
    with torch.no_grad():
        return _native.inverse(z, w, G, orders)
