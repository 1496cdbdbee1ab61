"""PaddedConv2d -- the reference's corner-padded invertible k x k convolution, on B200 kernels.

Same constructor, attributes (`.conv.weight`, `.mask`, `.pad`, `.order`, `.kernel_size`),
state-dict keys and return conventions as fastflow/layers/conv.py:21-221:
forward -> (z, logdet), reverse -> (x, 0) (a tuple, like the reference), logdet().
Everything runs through include/fincflow_b200.h with G = 1.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native
from ..ops import finc_conv, finc_inverse
from .flowlayer import FlowLayer

ORDERS = ("TL", "TR", "BL", "BR")


def init_finc_weight_(weight: torch.Tensor, order: str) -> torch.Tensor:
    """reference reset_parameters, layers/conv.py:63-79: N(0, 0.05^2); in TL orientation
    W[o,o,-1,-1] = 1, W[o,i>o,-1,-1] = 0; then flipped on dim 3 (TR), 2 (BL), both (BR)."""
    with torch.no_grad():
        nn.init.normal_(weight, mean=0.0, std=0.05)
        for c_out in range(weight.shape[0]):
            weight[c_out, c_out, -1, -1] = 1.0
            weight[c_out, c_out + 1:, -1, -1] = 0.0
        if order == "TR":
            weight.copy_(torch.flip(weight, [3]))
        elif order == "BL":
            weight.copy_(torch.flip(weight, [2]))
        elif order == "BR":
            weight.copy_(torch.flip(weight, [2, 3]))
    return weight


def finc_grad_mask(C: int, kernel_size, order: str) -> torch.Tensor:
    """reference get_mask, layers/conv.py:81-96 (CPU tensor attribute, kept for compatibility;
    the kernels apply the mask from `orders`, they never read this tensor)."""
    mask = torch.ones(C, C, *kernel_size)
    for c_out in range(C):
        mask[c_out, c_out:, -1, -1] = 0.0
    if order == "TR":
        mask = torch.flip(mask, [3])
    elif order == "BL":
        mask = torch.flip(mask, [2])
    elif order == "BR":
        mask = torch.flip(mask, [2, 3])
    return mask


class PaddedConv2d(FlowLayer):
    def __init__(self, in_channels, out_channels, kernel_size, bias=False, order="TL",
                 mask_in_backward=False, logdet_mode="float"):
        super().__init__()
        assert len(kernel_size) == 2
        assert order in ORDERS, "unknown order: {}".format(order)
        assert in_channels == out_channels, "FInC convolutions are square in channels"
        self.kernel_size = tuple(kernel_size)
        self.order = order
        K_H, K_W = self.kernel_size
        # (left, right, top, bottom), layers/conv.py:41-55 -- informational, nothing is padded
        self.pad = {"TL": (K_W - 1, 0, K_H - 1, 0), "TR": (0, K_W - 1, K_H - 1, 0),
                    "BL": (K_W - 1, 0, 0, K_H - 1), "BR": (0, K_W - 1, 0, K_H - 1)}[order]
        # parameter holder with the reference's name (`conv.weight`); bias is ignored exactly
        # as in the reference (layers/conv.py:60)
        self.conv = nn.Conv2d(in_channels, out_channels, self.kernel_size, bias=False)
        self.mask_in_backward = mask_in_backward
        self.logdet_mode = logdet_mode  # "float": python 0.0 like the reference; "tensor": fused [B] logdet
        self._orders = _native.pack_orders([order])
        self.reset_parameters()

    def reset_parameters(self):
        init_finc_weight_(self.conv.weight.data, self.order)
        self.mask = self.get_mask()

    def get_mask(self):
        return finc_grad_mask(self.conv.weight.shape[0], self.kernel_size, self.order)

    def reset_gradients(self):
        """layers/conv.py:98-99 without the H2D mask copy: one tiny kernel, in place."""
        g = self.conv.weight.grad
        if g is not None and not self.mask_in_backward:
            _native.apply_grad_mask_(g, 1, self._orders)

    def forward(self, x, context=None, compute_expensive=None):
        want = self.logdet_mode == "tensor"
        z, logdet = finc_conv(x, self.conv.weight, 1, self._orders, self.mask_in_backward, want)
        return z, (logdet if want else 0.0)

    def reverse(self, x, context=None, compute_expensive=None):
        return finc_inverse(x, self.conv.weight, 1, self._orders), 0

    def logdet(self, x, context=None):
        if self.logdet_mode == "tensor":
            return _native.logdet(self.conv.weight.detach(), x.shape[0], x.shape[2], x.shape[3], 1, self._orders)
        return 0.0
