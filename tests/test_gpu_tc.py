"""GPU parity tests of the tensor-core path (tcgen05 3xTF32 implicit GEMM, csrc/tc_igemm.cuh) through
the C ABI: channels-last convolutions against an fp64 PyTorch convolution, the fused Coupling layer
(fastflow/layers/coupling.py:44-105) against its PyTorch formulas in fp64 and fp32 (TF32 off).

Tolerances (max-norm relative, the metric of north_star): 3xTF32 convolution <= 2e-6 (measured
1e-7 .. 2e-7; cuDNN's own fp32 is at 2e-7 .. 5e-6 on these shapes), coupling output and
log-determinant <= 1e-5 (measured <= 7e-7), round trip <= 1e-4 max-abs; the optional single-pass
TF32 mode only has to be TF32-accurate (2e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


CONV_CASES = [
    # B, H, W, Cin, N, taps
    (2, 16, 16, 32, 32, 1), (2, 16, 16, 64, 256, 1), (2, 16, 16, 64, 32, 9), (3, 14, 14, 32, 64, 9),
    (5, 7, 7, 96, 128, 9), (9, 4, 4, 64, 64, 9), (2, 32, 32, 32, 32, 9), (8, 16, 16, 512, 512, 1),
    (1, 1, 1, 32, 32, 9), (3, 5, 9, 64, 96, 9), (2, 3, 130, 32, 32, 9),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("npass", [3, 1])
def test_tc_conv_matches_fp64_convolution(case, npass):
    from fincflow_b200 import _native

    B, H, W, Cin, N, taps = case
    k = 3 if taps == 9 else 1
    torch.manual_seed(B * 1000 + H * 100 + Cin)
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(N, Cin, k, k, device="cuda") / (Cin * taps) ** 0.5
    bias = torch.randn(N, device="cuda")
    wp = _native.tc_conv_prepare_weights(w, 0)
    flags = _native.FLAG_TF32_1PASS if npass == 1 else 0
    y = _native.tc_conv_nhwc(x.permute(0, 2, 3, 1).contiguous(), wp, bias, N, taps, relu=True, flags=flags)
    ref = F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)).permute(0, 2, 3, 1)
    assert _relmax(y, ref) <= (2e-3 if npass == 1 else 2e-6)


def test_tc_conv_backward_data_weights_and_relu_mask():
    """mode 2 (transposed + flipped weights) is the backward-data convolution; `relu_mask` zeroes where
    the saved activation was not positive"""
    from fincflow_b200 import _native

    torch.manual_seed(5)
    B, H, W, Cin, N = 3, 8, 8, 64, 32
    x = torch.randn(B, Cin, H, W, device="cuda", dtype=torch.float64, requires_grad=True)
    w = torch.randn(N, Cin, 3, 3, device="cuda", dtype=torch.float64) / 24
    pre = F.conv2d(x, w, padding=1)
    dy = torch.randn_like(pre)
    F.relu(pre).backward(dy)
    g = (dy * (pre > 0)).float().permute(0, 2, 3, 1).contiguous()     # gradient at the conv output, channels-last
    wt = _native.tc_conv_prepare_weights(w.float().contiguous(), 2)
    zero = torch.zeros(Cin, device="cuda")
    dx = _native.tc_conv_nhwc(g, wt, zero, Cin, 9)
    assert _relmax(dx, x.grad.permute(0, 2, 3, 1)) <= 2e-6
    # masked variant: dx * (m > 0) for an arbitrary channels-last mask tensor
    m = torch.randn(B, H, W, Cin, device="cuda")
    dxm = _native.tc_conv_nhwc(g, wt, zero, Cin, 9, relu_mask=m)
    assert torch.equal(dxm, dx * (m > 0))


COUPLING_CASES = [
    # B, C, H, W, width
    (4, 12, 16, 16, 512), (6, 24, 8, 8, 512), (9, 48, 4, 4, 512), (3, 4, 14, 14, 64), (3, 8, 7, 7, 64),
    (2, 96, 4, 4, 512), (2, 12, 32, 32, 128), (1, 2, 5, 3, 32),
]


def _coupling(C, H, W, width, seed=0):
    from fincflow_b200.flows import Coupling

    torch.manual_seed(seed)
    cp = Coupling((C, H, W), width=width).cuda()
    with torch.no_grad():  # the last conv is zero-initialised in the reference: give it something to do
        cp.net[4].weight.normal_(0, 0.02)
        cp.net[4].bias.normal_(0, 0.1)
        cp.net[4].logs.normal_(0, 0.1)
    return cp


@pytest.mark.parametrize("case", COUPLING_CASES)
def test_coupling_forward_reverse_match_reference_formulas(case):
    from fincflow_b200.flows import Coupling

    B, C, H, W, width = case
    cp = _coupling(C, H, W, width)
    x = torch.randn(B, C, H, W, device="cuda")
    with torch.no_grad():
        assert cp._use_tc(x)
        y, ld = cp(x)                       # tensor-core path
        xr = cp.reverse(y)
        Coupling.tensor_core = False
        try:
            y32, ld32 = cp(x)               # PyTorch formulas, fp32 (TF32 off)
            cpd = Coupling((C, H, W), width=width).cuda().double()
            cpd.load_state_dict({k: v.double() for k, v in cp.state_dict().items()})
            y64, ld64 = cpd(x.double())
        finally:
            Coupling.tensor_core = True
    assert torch.equal(y[:, :C // 2], x[:, :C // 2])
    assert _relmax(y, y64) <= 1e-5 and _relmax(ld, ld64) <= 1e-5
    assert _relmax(y, y32) <= 1e-5 and _relmax(ld, ld32) <= 1e-5
    assert float((xr - x).abs().max()) <= 1e-4


def test_coupling_single_pass_tf32_mode_and_weight_cache():
    from fincflow_b200.flows import Coupling

    cp = _coupling(12, 16, 16, 512, seed=3)
    x = torch.randn(5, 12, 16, 16, device="cuda")
    with torch.no_grad():
        y, ld = cp(x)
        blob = cp._blob
        cp(x)
        assert cp._blob is blob                       # no re-preparation while the weights are unchanged
        cp.net[2].bias.add_(0.25)                     # in-place update (what an optimizer does)
        y2, _ = cp(x)
        assert not torch.equal(y, y2)
        Coupling.tensor_core = False
        try:
            y2_ref, _ = cp(x)
        finally:
            Coupling.tensor_core = True
        assert _relmax(y2, y2_ref) <= 1e-5
        cp.precision = "tf32"
        y1, ld1 = cp(x)
        assert 1e-7 < _relmax(y1, y2_ref) <= 2e-3


def test_coupling_autograd_matches_pytorch_path():
    """training step through the fused forward: gradients equal the PyTorch-formula gradients"""
    from fincflow_b200.flows import Coupling

    cp = _coupling(12, 8, 8, 64, seed=7)
    x = torch.randn(4, 12, 8, 8, device="cuda", requires_grad=True)
    gy, gl = torch.randn(4, 12, 8, 8, device="cuda"), torch.randn(4, device="cuda")
    y, ld = cp(x)
    ((y * gy).sum() + (ld * gl).sum()).backward()
    got = [x.grad.clone()] + [p.grad.clone() for p in cp.parameters()]
    x.grad = None
    cp.zero_grad()
    Coupling.tensor_core = False
    try:
        y, ld = cp(x)
        ((y * gy).sum() + (ld * gl).sum()).backward()
    finally:
        Coupling.tensor_core = True
    want = [x.grad] + [p.grad for p in cp.parameters()]
    for a, b in zip(got, want):
        assert _relmax(a, b) <= 1e-5


@pytest.mark.parametrize("case", [(4096, 128, 64), (4096, 512, 128), (5000, 512, 160), (65536, 512, 512), (1000, 256, 320),
                                  (33, 128, 448)])
@pytest.mark.parametrize("npass", [3, 1])
def test_tc_wgrad_matches_fp64_matmul(case, npass):
    """dW[m, n] = sum_p P[p, m] Q[p, n] (reduction over pixels, split-K, fixed-order slice sum)"""
    from fincflow_b200 import _native

    npix, M, N = case
    torch.manual_seed(npix + M)
    P = torch.randn(npix, M, device="cuda")
    Q = torch.randn(npix, N, device="cuda")
    flags = _native.FLAG_TF32_1PASS if npass == 1 else 0
    dW = _native.tc_wgrad(P, Q, flags=flags)
    ref = P.double().t() @ Q.double()
    assert _relmax(dW, ref) <= (2e-3 if npass == 1 else 2e-6)
    assert torch.equal(dW, _native.tc_wgrad(P, Q, flags=flags))      # deterministic


@pytest.mark.parametrize("case", [(4, 12, 16, 16, 128), (3, 24, 8, 8, 256), (5, 48, 4, 4, 128), (2, 4, 14, 14, 128),
                                  (2, 96, 4, 4, 128), (8, 12, 16, 16, 512)])
def test_coupling_native_backward_matches_pytorch_autograd(case):
    """finc_coupling_backward_f32 (three weight-gradient GEMMs, three backward-data GEMMs, pointwise part,
    col2im) against autograd through the PyTorch formulas, fp32 with TF32 off"""
    from fincflow_b200 import _native
    from fincflow_b200.flows import Coupling

    B, C, H, W, width = case
    assert _native.coupling_prepared_bytes(C, width, True) > 0
    cp = _coupling(C, H, W, width, seed=11)
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    gy, gl = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, device="cuda")
    y, ld = cp(x)
    ((y * gy).sum() + (ld * gl).sum()).backward()
    got = [x.grad.clone()] + [p.grad.clone() for p in cp.parameters()]
    x.grad = None
    cp.zero_grad()
    Coupling.tensor_core = False
    try:
        y, ld = cp(x)
        ((y * gy).sum() + (ld * gl).sum()).backward()
    finally:
        Coupling.tensor_core = True
    want = [x.grad] + [p.grad for p in cp.parameters()]
    names = ["dx"] + [n for n, _ in cp.named_parameters()]
    for name, a, b in zip(names, got, want):
        assert a.shape == b.shape, name
        assert _relmax(a, b) <= 2e-5, (name, _relmax(a, b))
