"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a
GPU, exports every symbol include/fincflow_b200.h declares, and the host layer refuses to
run anywhere but on the CUDA kernels (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO


@pytest.fixture(scope="module")
def lib_path():
    from fincflow_b200 import build

    return build.build()


def _header_functions():
    src = open(os.path.join(REPO, "include", "fincflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(finc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _header_functions()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fincflow_b200.h but not exported"
    from fincflow_b200 import _native

    assert sorted(_native.SYMBOLS) == names


def test_abi_version_and_error_strings(lib_path):
    from fincflow_b200 import _native

    lib = _native.load()
    assert lib.finc_abi_version() == 1
    assert b"bad argument" in lib.finc_error_string(-1)
    assert b"workspace" in lib.finc_error_string(-2)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.finc_forward_f32(None, None, None, None, 1, 4, 3, 8, 8, 3, 3, 0xE4, 0, None) == -1
    assert lib.finc_inverse_f32(None, None, None, 1, 17, 3, 8, 8, 3, 3, 0, 0, None) == -1  # G > 16
    assert lib.finc_backward_weight_workspace_bytes(256, 4, 3, 16, 16, 3, 3) >= 4096
    # tensor-core entry points: sizes are host arithmetic, bad arguments are rejected before any CUDA call
    assert lib.finc_coupling_prepared_bytes(12, 512, 0) > 4 * (2 * 512 * 64 + 2 * 512 * 512 + 2 * 9 * 16 * 512)
    assert lib.finc_coupling_prepared_bytes(12, 512, 1) > lib.finc_coupling_prepared_bytes(12, 512, 0)
    assert lib.finc_coupling_prepared_bytes(12, 16, 0) == 0 and lib.finc_coupling_prepared_bytes(13, 512, 0) == 0
    assert lib.finc_coupling_prepared_bytes(12, 96, 0) > 0 and lib.finc_coupling_prepared_bytes(12, 96, 1) == 0
    assert lib.finc_coupling_backward_workspace_bytes(256, 12, 16, 16, 512) > 0
    assert lib.finc_tc_wgrad_workspace_bytes(65536, 512, 512) > 0 and lib.finc_tc_wgrad_workspace_bytes(100, 512, 100) == 0
    assert lib.finc_slogdet_inverse_f32(None, None, None, 3, 200, None) == -1
    assert lib.finc_coupling_workspace_bytes(256, 12, 16, 16, 512) >= 4 * 256 * 256 * (64 + 512 + 512)
    assert lib.finc_coupling_apply_f32(None, None, None, None, None, 0, 1, 12, 8, 8, 512, 0, 0, None) == -1
    assert lib.finc_tc_conv_nhwc_f32(None, None, None, None, None, 1, 8, 8, 32, 32, 9, 0, 0, None) == -1
    assert lib.finc_tc_conv_weights_bytes(512, 6, 9, 1) == 2 * 512 * 64 * 4


def test_sass_is_sm100a_with_bulk_tma(lib_path):
    """the tiled kernels must really be Blackwell code using the TMA bulk-copy engine"""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out  # cp.async.bulk global<->shared
    assert "SYNCS" in out   # mbarrier
    # tensor-core path: tcgen05.mma / tcgen05.ld / TMA tensor-map loads and stores
    # (/opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"):
        assert mnemonic in out, mnemonic
    assert "HMMA.16" not in out and "HGMMA" not in out   # no legacy mma.sync / wgmma paths


def test_no_cpu_fallback():
    from fincflow_b200 import _native
    from fincflow_b200.fastflow import FastFlowUnit
    from fincflow_b200.layers.conv import PaddedConv2d

    with pytest.raises(_native.FincNativeError):
        FastFlowUnit(8, 8, (3, 3))(torch.randn(2, 8, 4, 4))
    with pytest.raises(_native.FincNativeError):
        PaddedConv2d(3, 3, (3, 3), order="BR").reverse(torch.randn(2, 3, 4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "fincflow_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.lower() or f == "build.py", f"{f} mentions the oracle"


def test_layer_api_matches_reference_surface():
    """names/attributes a reference user relies on"""
    from fincflow_b200.fastflow import FastFlowUnit
    from fincflow_b200.layers.conv import PaddedConv2d
    from fincflow_b200.layers.flowlayer import FlowLayer

    conv = PaddedConv2d(3, 3, (3, 3), order="TR")
    assert isinstance(conv, FlowLayer)
    assert conv.pad == (0, 2, 2, 0) and conv.order == "TR" and conv.kernel_size == (3, 3)
    assert list(conv.state_dict().keys()) == ["conv.weight"]
    w = conv.conv.weight.data
    # stored orientation of TR: corner tap (kH-1, 0): unit diagonal, zero above it
    for o in range(3):
        assert w[o, o, 2, 0] == 1.0 and (w[o, o + 1:, 2, 0] == 0).all()
    assert conv.mask.sum() == conv.mask.numel() - 3 * 4 // 2
    unit = FastFlowUnit(12, 12, (3, 3))
    assert sorted(unit.state_dict().keys()) == sorted(
        f"conv_{q}.conv.weight" for q in ("tl", "tr", "bl", "br"))
    other = FastFlowUnit(12, 12, 3)
    other.load_state_dict(unit.state_dict())
    assert torch.equal(other.weight, unit.weight)
    assert unit.conv_br.conv.weight.shape == (3, 3, 3, 3) and unit.conv_br.order == "BR"
    with pytest.raises(AssertionError):
        FastFlowUnit(6, 6, (3, 3))
