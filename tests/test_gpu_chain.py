"""finc_chain_f32 (a chain of FInC units, optionally with the ActNorm o Conv1x1 affine map after each, in one
launch) against the per-unit launches it replaces (fastflow/fastflow.py:31-50, layers/actnorm.py:14-52,
layers/conv1x1.py:18-43) and against the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err

REL_TOL = 1e-5
from oracle import finc_oracle as fo

pytestmark = pytest.mark.gpu

SHAPES = [  # B, C (= 4*Cq), H, W, units
    (256, 12, 16, 16, 16), (256, 24, 8, 8, 16), (256, 48, 4, 4, 16),       # cfg3 levels
    (37, 12, 16, 16, 3), (5, 48, 4, 4, 2), (1, 24, 8, 8, 1),               # ragged / tiny batches
    (128, 4, 14, 14, 16), (128, 8, 7, 7, 1),                               # cfg2 (MNIST) levels: WT = 2 and 1
    (64, 12, 32, 32, 4), (40, 96, 4, 4, 3), (9, 96, 8, 8, 2),              # cfg4 / cfg5 shapes (Cq = 3, 24)
]


def _units(U, C, seed):
    from fincflow_b200.fastflow import FastFlowUnit

    torch.manual_seed(seed)
    return torch.stack([FastFlowUnit(C, C, (3, 3)).weight.detach() for _ in range(U)]).cuda().contiguous()


@pytest.mark.parametrize("shape", SHAPES)
def test_forward_chain_is_bit_identical_to_per_unit_launches(shape):
    from fincflow_b200 import _native

    B, C, H, W, U = shape
    assert _native.chain_supported(4, C // 4, H, W)
    w = _units(U, C, B + C)
    x = torch.randn(B, C, H, W, device="cuda")
    out = torch.full((U, B, C, H, W), float("nan"), device="cuda")
    ld = torch.zeros(B, device="cuda")
    _native.chain(x, w, out, logdet_out=ld)
    cur, ld_ref = x, torch.zeros(B, device="cuda")
    for u in range(U):
        cur, l = _native.forward(cur, w[u])
        ld_ref += l
        assert torch.equal(out[u], cur), f"unit {u}"
    assert rel_err(ld.cpu().numpy(), ld_ref.cpu().numpy()) <= 1e-6 or float(ld_ref.abs().max()) == 0.0
    # only the last result
    last = torch.empty_like(x)
    _native.chain(x, w, last)
    assert torch.equal(last, out[U - 1])
    if B <= 40:
        z = fo.forward(x.cpu().numpy(), w[0].cpu().numpy())
        assert rel_err(out[0].cpu().numpy(), z) <= REL_TOL


@pytest.mark.parametrize("shape", SHAPES[:8])
def test_backward_data_chain(shape):
    from fincflow_b200 import _native

    B, C, H, W, U = shape
    w = _units(U, C, 3 * B + C)
    dz = torch.randn(B, C, H, W, device="cuda")
    out = torch.full((U, B, C, H, W), float("nan"), device="cuda")
    # dzs[u] = FInC_u^T(dzs[u + 1]) for u = U-1 .. 0
    _native.chain(dz, w, out, units=range(U - 1, -1, -1), transpose=True)
    cur = dz
    for u in reversed(range(U)):
        cur = _native.backward_input(cur, w[u])
        assert torch.equal(out[u], cur), f"unit {u}"
    # a partial chain (the trainer never needs the data gradient of unit 0)
    if U > 1:
        out2 = torch.full((U, B, C, H, W), float("nan"), device="cuda")
        _native.chain(dz, w, out2, units=range(U - 1, 0, -1), transpose=True)
        assert torch.equal(out2[1:], out[1:]) and bool(torch.isnan(out2[0]).all())


@pytest.mark.parametrize("shape", [(256, 12, 16, 16, 1), (256, 24, 8, 8, 1), (256, 48, 4, 4, 1), (19, 48, 4, 4, 3),
                                   (128, 4, 14, 14, 2), (128, 8, 7, 7, 1)])
def test_chain_with_affine_glue(shape):
    """FastFlowUnit + ActNorm + Conv1x1 in one launch == the FInC launch followed by finc_affine1x1_f32"""
    from fincflow_b200 import _native

    B, C, H, W, U = shape
    w = _units(U, C, 7 * B + C)
    torch.manual_seed(B)
    A = (torch.linalg.qr(torch.randn(U, C, C))[0] * torch.exp(0.1 * torch.randn(U, 1, C))).cuda().contiguous()
    b = torch.randn(U, C).cuda()
    x = torch.randn(B, C, H, W, device="cuda")
    out = torch.empty(U, B, C, H, W, device="cuda")
    ld = torch.zeros(B, device="cuda")
    _native.chain(x, w, out, A=A, bias=b, logdet_out=ld)
    cur = x
    for u in range(U):
        z, _ = _native.forward(cur, w[u])
        cur = _native.affine1x1(z, A[u], b[u])
        assert rel_err(out[u].cpu().numpy(), cur.cpu().numpy()) <= 2e-6, f"unit {u}"
    want = (A[0].double() @ torch.from_numpy(fo.forward(x[:4].cpu().numpy(), w[0].cpu().numpy())).cuda().flatten(2)
            + b[0].double()[:, None]).view(4, C, H, W)
    assert rel_err(out[0, :4].cpu().numpy(), want.cpu().numpy()) <= REL_TOL


def test_chain_rejects_what_it_does_not_cover():
    from fincflow_b200 import _native

    assert not _native.chain_supported(4, 3, 16, 16, (5, 5))
    assert not _native.chain_supported(4, 5, 16, 16)
    w = _units(2, 12, 0)
    x = torch.randn(4, 12, 8, 8, device="cuda")
    with pytest.raises(_native.FincNativeError):
        _native.chain(x, w, torch.empty(2, 4, 12, 8, 8, device="cuda"), units=[0, 0])
    with pytest.raises(_native.FincNativeError):
        _native.chain(x, w, torch.empty(3, 4, 12, 8, 8, device="cuda"))


@pytest.mark.parametrize("shape", [(256, 12, 16, 16, 16, 3), (256, 24, 8, 8, 16, 3), (256, 48, 4, 4, 16, 3), (37, 12, 16, 16, 3, 3),
                                   (5, 48, 4, 4, 2, 3), (128, 4, 14, 14, 16, 3), (64, 12, 32, 32, 4, 3), (40, 12, 16, 16, 3, 5),
                                   (1024, 12, 16, 16, 8, 3)])
def test_inverse_chain_is_bit_identical_to_per_unit_launches(shape):
    from fincflow_b200 import _native

    B, C, H, W, U, k = shape
    from fincflow_b200.fastflow import FastFlowUnit

    torch.manual_seed(B + C + U)
    w = torch.stack([FastFlowUnit(C, C, (k, k)).weight.detach() for _ in range(U)]).cuda().contiguous()
    nb = _native.prepared_weights_bytes(_native.PREP_INVERSE, B, 4, C // 4, H, W, k, k)
    assert nb > 0
    tables = torch.empty((U, nb), dtype=torch.uint8, device="cuda")
    _native.prepare_weights(w, tables, _native.PREP_INVERSE, B, H, W)
    z = torch.randn(B, C, H, W, device="cuda")
    x = _native.inverse_chain(z, tables, (k, k), range(U - 1, -1, -1))
    cur = z
    for u in reversed(range(U)):
        cur = _native.inverse(cur, w[u])
    assert torch.equal(x, cur)
    # a sub-chain, ascending order
    if U >= 3:
        x2 = _native.inverse_chain(z, tables, (k, k), range(0, 2))
        assert torch.equal(x2, _native.inverse(_native.inverse(z, w[0]), w[1]))
    if B <= 40:
        zz = x.cpu().numpy()
        for u in range(U):
            zz = fo.forward(zz, w[u].cpu().numpy())
        assert np.abs(zz - z.cpu().numpy()).max() <= 1e-4


@pytest.mark.parametrize("shape", [(256, 12, 16, 16, 15, 3), (256, 24, 8, 8, 15, 3), (256, 48, 4, 4, 15, 3), (37, 12, 16, 16, 3, 3),
                                   (128, 4, 14, 14, 16, 3), (64, 12, 32, 32, 4, 3), (40, 12, 16, 16, 2, 5), (5, 48, 4, 4, 1, 3)])
def test_batched_weight_gradient_equals_per_unit_launches(shape):
    from fincflow_b200 import _native

    B, C, H, W, U, k = shape
    torch.manual_seed(B * U + C)
    dz = torch.randn(U + 1, B, C, H, W, device="cuda")
    x = torch.randn(U + 1, B, C, H, W, device="cuda")
    out = torch.full((U, C, C // 4, k, k), float("nan"), device="cuda")
    # unit u reads dz[u + 1] and x[u]: the slices the trainer passes
    _native.backward_weight_batched(dz[1:], x[:U], out, (k, k))
    for u in range(U):
        want = _native.backward_weight(dz[u + 1], x[u], (k, k))
        assert rel_err(out[u].cpu().numpy(), want.cpu().numpy()) <= 2e-6, f"unit {u}"
        assert torch.equal(out[u] == 0, want == 0)               # the same masked entries
    if B <= 40:
        ref = fo.backward_weight(dz[1].cpu().numpy(), x[0].cpu().numpy(), (k, k))
        assert rel_err(out[0].cpu().numpy(), ref) <= REL_TOL
    # twice on the same workspace (tickets must be left clean)
    ws = torch.zeros(_native.backward_weight_batched_workspace_bytes(B, 4, C // 4, H, W, k, k, U), dtype=torch.uint8, device="cuda")
    a = _native.backward_weight_batched(dz[1:], x[:U], torch.empty_like(out), (k, k), workspace=ws)
    b = _native.backward_weight_batched(dz[1:], x[:U], torch.empty_like(out), (k, k), workspace=ws)
    assert torch.equal(a, b) and torch.equal(a, out)
