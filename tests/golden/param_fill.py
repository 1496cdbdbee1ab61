"""Deterministic parameters for whole-flow fixtures at the BASELINE configurations.

The full-depth models have tens of millions of parameters (CIFAR-10 flow: 22.9 M), far too many to
commit.  Instead both sides -- the reference model in tests/golden/make_golden_flow_full.py and our model
in tests/test_gpu_flow.py -- fill their state dict with THIS function: every tensor is a function of its
state-dict key and shape only (generator seeded with crc32(key)), independent of construction order.
"""
import zlib

import torch

_FLIP = {"tl": [], "tr": [3], "bl": [2], "br": [2, 3]}


def _gen(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ seed) & 0x7FFFFFFF)


def filled_state_dict(model, seed=0):
    """{key: tensor} for every entry of model.state_dict() this scheme covers (reference key names)"""
    out = {}
    for key, ref in model.state_dict().items():
        g = _gen(key, seed)
        shape = tuple(ref.shape)
        if ".fastflow_unit.conv_" in key and key.endswith(".conv.weight"):
            q = key.split(".fastflow_unit.conv_")[1][:2]
            w = torch.randn(shape, generator=g) * 0.02        # layers/conv.py:63-79 with a seeded draw (std 0.02: 48 steps stay O(1))
            for o in range(shape[0]):
                w[o, o, -1, -1] = 1.0
                w[o, o + 1:, -1, -1] = 0.0
            out[key] = torch.flip(w, _FLIP[q]).contiguous() if _FLIP[q] else w
        elif key.endswith("actnorm.translation") or key.endswith("actnorm.log_scale"):
            out[key] = torch.randn(shape, generator=g) * 0.05
        elif key.endswith("actnorm.initialized"):
            out[key] = torch.ones_like(ref)
        elif key.endswith("conv1x1.W"):
            out[key] = torch.linalg.qr(torch.randn(shape, generator=g))[0].contiguous()
        elif ".net." in key and key.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            std = 0.05 / fan_in ** 0.5 if key.endswith("net.4.weight") else 1.0 / fan_in ** 0.5
            out[key] = torch.randn(shape, generator=g) * std
        elif key.endswith("net.4.bias") or key.endswith("net.4.logs"):
            # reference quirk: Conv2dZero builds `bias` and `logs` from the SAME zeros tensor
            # (layers/coupling.py:31-35), so the two parameters share storage and are always equal
            g = _gen(key[:-len("bias")] + "logs" if key.endswith("bias") else key, seed)
            out[key] = torch.randn(shape, generator=g) * 0.02
        elif ".net." in key and key.endswith(".bias"):
            out[key] = torch.randn(shape, generator=g) * 0.05
    return out
