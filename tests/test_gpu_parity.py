"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the
C ABI (fincflow_b200._native -> libfincflow_b200.so); the oracle is only the checker.

Tolerances (BASELINE.json north_star): fp32, <= 1e-5 relative on z and logdet
(max-norm relative: max|a-b| / max|b|), <= 1e-4 max-abs round trip
||x - reverse(forward(x))||.  dX / dW / inverse get the same 1e-5 relative bar.
"""
import numpy as np
import pytest
import torch

from conftest import elementwise_close, Golden, rel_err
from oracle import finc_oracle as fo

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
RT_TOL = 1e-4

_G = Golden()
FULL = [n for n in _G.names if not n.startswith("kat_")]
KATS = [n for n in _G.names if n.startswith("kat_")]


@pytest.fixture(scope="module")
def nat():
    from fincflow_b200 import _native

    _native.load()
    return _native


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def run_all(nat, x, w, dz, zs, orders, flags=0):
    G = len(orders)
    po = nat.pack_orders(orders)
    k = tuple(w.shape[2:])
    xd, wd, dzd, zsd = dev(x), dev(w), dev(dz), dev(zs)
    z, logdet = nat.forward(xd, wd, G, po, True, flags)
    dx = nat.backward_input(dzd, wd, G, po, flags)
    dw = nat.backward_weight(dzd, xd, k, G, po, flags)
    dw_raw = nat.backward_weight(dzd, xd, k, G, po, flags | nat.FLAG_NO_MASK)
    x_zs = nat.inverse(zsd, wd, G, po, flags)
    x_rt = nat.inverse(z, wd, G, po, flags)
    torch.cuda.synchronize()
    return {k_: v.cpu().numpy() for k_, v in dict(z=z, logdet=logdet, dx=dx, dw=dw, dw_raw=dw_raw,
                                                  x_zs=x_zs, x_rt=x_rt).items()}


@pytest.mark.parametrize("flags", [0, 1], ids=["tiled", "naive"])
@pytest.mark.parametrize("name", FULL)
def test_golden_vectors_from_reference(nat, golden, name, flags):
    """outputs of the reference's own PaddedConv2d / FastFlowUnit / Cython solver"""
    c = golden.case(name)
    orders = tuple(int(o) for o in c["orders"])
    r = run_all(nat, c["x"], c["w"], c["dz"], c["zs"], orders, flags)
    assert rel_err(r["z"], c["z"]) <= REL_TOL
    # element-wise (with absolute floor) next to the max-norm metric, on every output
    for got, want in (("z", "z"), ("dx", "dx"), ("dw_raw", "dw_raw"), ("dw", "dw_masked"), ("x_zs", "x_from_zs")):
        assert elementwise_close(r[got], c[want]), (got, want)
    assert np.abs(r["logdet"]).max() == 0.0  # reference: python float 0.0
    assert rel_err(r["dx"], c["dx"]) <= REL_TOL
    assert rel_err(r["dw_raw"], c["dw_raw"]) <= REL_TOL
    assert rel_err(r["dw"], c["dw_masked"]) <= REL_TOL
    assert np.array_equal(r["dw"] == 0, c["dw_masked"] == 0)
    assert rel_err(r["x_zs"], c["x_from_zs"]) <= REL_TOL
    assert np.abs(r["x_rt"] - c["x"]).max() <= RT_TOL


@pytest.mark.parametrize("flags", [0, 1], ids=["tiled", "naive"])
@pytest.mark.parametrize("name", KATS)
def test_known_answer_vectors_exact(nat, golden, name, flags):
    c = golden.case(name)
    orders = tuple(int(o) for o in c["orders"])
    G, po = len(orders), nat.pack_orders(orders)
    x = nat.inverse(dev(c["zs"]), dev(c["w"]), G, po, flags).cpu().numpy()
    assert np.array_equal(x, c["x_from_zs"])  # small integers: bit exact
    z, _ = nat.forward(dev(c["x_from_zs"]), dev(c["w"]), G, po, False, flags)
    assert np.array_equal(z.cpu().numpy(), c["zs"])


# (B, G, C, H, W, kH, kW, orders): BASELINE configs' unit tensors at small batch + edge cases
ORACLE_CASES = [
    (64, 1, 4, 14, 14, 3, 3, (0,)),           # cfg1 single TL conv C=4
    (64, 4, 1, 14, 14, 3, 3, (0, 1, 2, 3)),   # cfg1 as FastFlowUnit(4)
    (33, 4, 2, 7, 7, 3, 3, (0, 1, 2, 3)),     # MNIST final level (tile not 16B multiple)
    (37, 4, 3, 16, 16, 3, 3, (0, 1, 2, 3)),   # CIFAR L0, ragged batch
    (19, 4, 6, 8, 8, 3, 3, (0, 1, 2, 3)),
    (21, 4, 12, 4, 4, 3, 3, (0, 1, 2, 3)),
    (5, 4, 3, 32, 32, 3, 3, (0, 1, 2, 3)),    # ImageNet64 L0
    (5, 4, 24, 4, 4, 3, 3, (0, 1, 2, 3)),
    (3, 4, 3, 32, 32, 5, 5, (0, 1, 2, 3)),    # k=5
    (3, 4, 24, 4, 4, 5, 5, (0, 1, 2, 3)),
    (1, 4, 3, 16, 16, 3, 3, (0, 1, 2, 3)),    # sampling n=1
    (2, 1, 3, 9, 5, 3, 3, (3,)),              # H > W (reference's Cython solver is wrong here)
    (2, 2, 3, 12, 40, 3, 3, (1, 2)),          # W > 32: two column blocks
    (2, 1, 5, 6, 6, 2, 3, (2,)),              # non-square kernel -> generic kernels
    (2, 1, 10, 28, 28, 3, 3, (0,)),           # reference test shape C=10 (test_examples.py:218-222)
    (1, 1, 50, 8, 8, 3, 3, (1,)),             # C=50 -> generic inverse
    (2, 4, 3, 1, 1, 3, 3, (0, 1, 2, 3)),      # 1x1 image
    (2, 3, 2, 3, 2, 5, 5, (3, 0, 1)),         # kernel larger than the image, G=3
    (300, 4, 3, 8, 8, 3, 3, (0, 1, 2, 3)),    # many items per warp: exercises the pipeline
    (700, 4, 3, 7, 8, 3, 3, (0, 1, 2, 3)),    # odd H, batch large enough for the row-blocked conv variant
    (900, 4, 2, 5, 4, 3, 3, (3, 2, 1, 0)),
    (1500, 4, 1, 3, 12, 3, 3, (0, 1, 2, 3)),
    (6000, 4, 3, 16, 16, 3, 3, (0, 1, 2, 3)),  # CIFAR L0 at a batch that selects it (2 rows per sub-item)
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=lambda c: "B{}G{}C{}_{}x{}_k{}x{}".format(*c[:7]))
def test_against_oracle(nat, case):
    B, G, C, H, W, kH, kW, orders = case
    rng = np.random.default_rng(hash(case[:7]) % (2**32))
    w = np.concatenate([fo.init_weight(C, (kH, kW), o, rng) for o in orders], 0)
    x = rng.normal(size=(B, G * C, H, W)).astype(np.float32)
    dz = rng.normal(size=x.shape).astype(np.float32)
    zs = rng.normal(size=x.shape).astype(np.float32)
    r = run_all(nat, x, w, dz, zs, orders)
    assert rel_err(r["z"], fo.forward(x, w, orders)) <= REL_TOL
    assert rel_err(r["dx"], fo.backward_input(dz, w, orders)) <= REL_TOL
    assert rel_err(r["dw"], fo.backward_weight(dz, x, (kH, kW), orders)) <= REL_TOL
    assert rel_err(r["dw_raw"], fo.backward_weight(dz, x, (kH, kW), orders, apply_mask=False)) <= REL_TOL
    assert rel_err(r["x_zs"], fo.inverse(zs, w, orders)) <= REL_TOL
    assert elementwise_close(r["z"], fo.forward(x, w, orders)) and elementwise_close(r["x_zs"], fo.inverse(zs, w, orders))
    assert elementwise_close(r["dx"], fo.backward_input(dz, w, orders))
    assert elementwise_close(r["dw"], fo.backward_weight(dz, x, (kH, kW), orders))
    assert np.abs(r["x_rt"] - x).max() <= RT_TOL
    assert np.abs(r["logdet"]).max() == 0.0


def test_logdet_general_formula(nat):
    """diagonal != 1: logdet = H*W*sum log|diag|"""
    rng = np.random.default_rng(2)
    w = fo.init_unit_weight(3, (3, 3), rng)
    w[0, 0, 2, 2] = 2.0      # TL corner (kH-1, kW-1)
    w[3 + 1, 1, 2, 0] = -0.5  # TR corner (kH-1, 0)
    x = rng.normal(size=(7, 12, 8, 8)).astype(np.float32)
    _, ld = nat.forward(dev(x), dev(w), 4, nat.ORDERS_UNIT, True)
    want = fo.logdet(w, 8, 8)
    assert abs(want - 64 * (np.log(2.0) + np.log(0.5))) < 1e-9
    assert rel_err(ld.cpu().numpy(), np.full(7, want)) <= REL_TOL or abs(want) < 1e-6
    ld2 = nat.logdet(dev(w), 7, 8, 8)
    assert torch.allclose(ld, ld2, rtol=1e-6, atol=1e-6)


def test_unaligned_and_noncontiguous_inputs(nat):
    rng = np.random.default_rng(9)
    w = fo.init_unit_weight(3, (3, 3), rng)
    x = rng.normal(size=(4, 12, 8, 8)).astype(np.float32)
    big = torch.zeros(x.size + 1, device="cuda")
    xv = big[1:].view(4, 12, 8, 8)  # 4-byte aligned only -> non-bulk copy path
    xv.copy_(dev(x))
    z, _ = nat.forward(xv, dev(w), 4, nat.ORDERS_UNIT, False)
    assert rel_err(z.cpu().numpy(), fo.forward(x, w)) <= REL_TOL
    xi = nat.inverse(xv, dev(w), 4, nat.ORDERS_UNIT)
    assert rel_err(xi.cpu().numpy(), fo.inverse(x, w)) <= REL_TOL
    dw = nat.backward_weight(xv, xv, (3, 3), 4, nat.ORDERS_UNIT)
    assert rel_err(dw.cpu().numpy(), fo.backward_weight(x, x, (3, 3))) <= REL_TOL
    # channels-last strided view: host layer makes it contiguous
    xt = dev(x).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    z2, _ = nat.forward(xt, dev(w), 4, nat.ORDERS_UNIT, False)
    assert torch.equal(z2, nat.forward(dev(x), dev(w), 4, nat.ORDERS_UNIT, False)[0])


def test_inverse_in_place_and_empty_batch(nat):
    rng = np.random.default_rng(4)
    w = dev(fo.init_unit_weight(6, (3, 3), rng))
    z = dev(rng.normal(size=(9, 24, 8, 8)))
    want = nat.inverse(z, w)
    buf = z.clone()
    nat.inverse(buf, w, out=buf)
    assert torch.equal(buf, want)
    e = torch.empty(0, 24, 8, 8, device="cuda")
    assert nat.forward(e, w)[0].shape == (0, 24, 8, 8)
    assert nat.inverse(e, w).shape == (0, 24, 8, 8)


@pytest.mark.parametrize("shape", [
    (128, 4, 14, 14, 3), (128, 8, 7, 7, 3),                                   # cfg2 MNIST, B=128
    (256, 12, 16, 16, 3), (256, 24, 8, 8, 3), (256, 48, 4, 4, 3),             # cfg3 CIFAR, B=256
    (512, 12, 16, 16, 3), (512, 24, 8, 8, 3), (512, 48, 4, 4, 3),             # cfg4 ImageNet32, B=512
    (64, 12, 32, 32, 3), (64, 24, 16, 16, 3), (64, 48, 8, 8, 3), (64, 96, 4, 4, 3),  # cfg5 k=3
    (64, 12, 32, 32, 5), (64, 24, 16, 16, 5), (64, 48, 8, 8, 5), (64, 96, 4, 4, 5),  # cfg5 k=5
], ids=lambda s: "B{}C{}_{}x{}_k{}".format(*s))
def test_full_size_properties(nat, shape):
    """BASELINE.json sizes through size-independent properties + the (fast) C oracle."""
    B, CT, H, W, k = shape
    torch.manual_seed(B * 1000 + CT + k)
    from fincflow_b200.fastflow import FastFlowUnit

    unit = FastFlowUnit(CT, CT, (k, k)).cuda()
    w = unit.weight.detach()
    x = torch.randn(B, CT, H, W, device="cuda")
    y = torch.randn(B, CT, H, W, device="cuda")
    z, ld = nat.forward(x, w)
    # round trip both ways
    assert (nat.inverse(z, w) - x).abs().max().item() <= RT_TOL
    zs = torch.randn_like(x)
    assert rel_err(nat.forward(nat.inverse(zs, w), w)[0].cpu().numpy(), zs.cpu().numpy()) <= REL_TOL
    assert ld.abs().max().item() == 0.0
    # linearity of forward and adjointness <dz, A x> == <A^T dz, x>, <dz, A x> == <dW, W> (A linear in W)
    zy = nat.forward(y, w)[0]
    zxy = nat.forward(x + 2 * y, w)[0]
    assert rel_err(zxy.cpu().numpy(), (z + 2 * zy).cpu().numpy()) <= REL_TOL
    dz = torch.randn_like(x)
    dx = nat.backward_input(dz, w)
    lhs = (dz.double() * z.double()).sum().item()
    assert abs(lhs - (dx.double() * x.double()).sum().item()) <= 1e-5 * (dz.norm() * z.norm()).item()
    dw_raw = nat.backward_weight(dz, x, (k, k), flags=nat.FLAG_NO_MASK)
    assert abs(lhs - (dw_raw.double() * w.double()).sum().item()) <= 1e-5 * (dz.norm() * z.norm()).item()
    # tiled kernels == generic kernels == oracle
    assert rel_err(z.cpu().numpy(), nat.forward(x, w, flags=nat.FLAG_NAIVE)[0].cpu().numpy()) <= REL_TOL
    assert rel_err(nat.inverse(zs, w).cpu().numpy(),
                   nat.inverse(zs, w, flags=nat.FLAG_NAIVE).cpu().numpy()) <= REL_TOL
    xn, wn = x.cpu().numpy(), w.cpu().numpy()
    assert rel_err(z.cpu().numpy(), fo.forward(xn, wn)) <= REL_TOL
    assert rel_err(nat.inverse(zs, w).cpu().numpy(), fo.inverse(zs.cpu().numpy(), wn)) <= REL_TOL
    assert rel_err(dx.cpu().numpy(), fo.backward_input(dz.cpu().numpy(), wn)) <= REL_TOL
    dw = nat.backward_weight(dz, x, (k, k))
    assert rel_err(dw.cpu().numpy(), fo.backward_weight(dz.cpu().numpy(), xn, (k, k))) <= REL_TOL
    # determinism: two launches give bit-identical results (no floating-point atomics)
    assert torch.equal(dw, nat.backward_weight(dz, x, (k, k)))


INVERSE_KERNEL_CASES = [
    # (B, CT, H, W, k): every (C, k, P, stacks-per-warp) family of the register-window kernel,
    # ragged batches (last item has partly empty stacks) and batches deep enough for stacked tiles
    (256, 12, 16, 16, 3), (256, 24, 8, 8, 3), (256, 48, 4, 4, 3), (1, 48, 4, 4, 3), (7, 24, 8, 8, 3),
    (3001, 12, 16, 16, 3), (5003, 24, 8, 8, 3), (9001, 48, 4, 4, 3), (20011, 48, 4, 4, 3),
    (130, 4, 14, 14, 3), (1500, 4, 14, 14, 3), (129, 8, 7, 7, 3), (1027, 8, 7, 7, 3),
    (33, 12, 32, 32, 3), (67, 16, 6, 10, 3), (64, 96, 4, 4, 3), (1031, 96, 4, 4, 3), (41, 96, 8, 8, 3),
    (35, 12, 32, 32, 5), (66, 24, 16, 16, 5), (67, 48, 8, 8, 5), (300, 48, 4, 4, 5), (64, 96, 4, 4, 5),
    (19, 4, 14, 14, 5), (23, 8, 7, 7, 5), (21, 16, 5, 3, 5), (2100, 12, 16, 16, 5),
]


@pytest.mark.parametrize("shape", INVERSE_KERNEL_CASES, ids=lambda s: "B{}C{}_{}x{}_k{}".format(*s))
def test_inverse_kernel_generations_agree(nat, shape):
    """register-window kernel (default) == shared-memory wavefront kernel == generic tiled kernel
    == oracle (on a slice), and inverts the forward at full size"""
    B, CT, H, W, k = shape
    torch.manual_seed(B + CT + k)
    from fincflow_b200.fastflow import FastFlowUnit

    w = FastFlowUnit(CT, CT, (k, k)).weight.detach().cuda()
    zs = torch.randn(B, CT, H, W, device="cuda")
    x_rw = nat.inverse(zs, w)
    x_wave = nat.inverse(zs, w, flags=nat.FLAG_WAVE_SMEM)
    x_gen = nat.inverse(zs, w, flags=nat.FLAG_GENERIC_TILED)
    assert rel_err(x_rw.cpu().numpy(), x_wave.cpu().numpy()) <= REL_TOL
    assert rel_err(x_rw.cpu().numpy(), x_gen.cpu().numpy()) <= REL_TOL
    assert rel_err(nat.forward(x_rw, w)[0].cpu().numpy(), zs.cpu().numpy()) <= REL_TOL
    # oracle on the first and last images (the last item is the ragged one)
    sel = torch.cat([zs[:3], zs[-3:]]) if B > 6 else zs
    got = torch.cat([x_rw[:3], x_rw[-3:]]) if B > 6 else x_rw
    assert rel_err(got.cpu().numpy(), fo.inverse(sel.cpu().numpy(), w.cpu().numpy())) <= REL_TOL
    # in place
    buf = zs.clone()
    nat.inverse(buf, w, out=buf)
    assert torch.equal(buf, x_rw)


def test_large_batch_beyond_reference_limit(nat):
    """B > 1024: the reference kernel maps the batch to blockDim.x and cannot launch
    (cinc_cuda_kernel_level2.cu:110-111)."""
    torch.manual_seed(0)
    from fincflow_b200.fastflow import FastFlowUnit

    unit = FastFlowUnit(12, 12, (3, 3)).cuda()
    x = torch.randn(2500, 12, 16, 16, device="cuda")
    z, _ = nat.forward(x, unit.weight.detach())
    assert (nat.inverse(z, unit.weight.detach()) - x).abs().max().item() <= RT_TOL
    assert rel_err(z[-3:].cpu().numpy(), fo.forward(x[-3:].cpu().numpy(), unit.weight.detach().cpu().numpy())) <= REL_TOL


def test_layers_autograd_and_reference_conventions(nat):
    """PaddedConv2d / FastFlowUnit through torch.autograd, reference return conventions."""
    from fincflow_b200.fastflow import FastFlowUnit, clear_grad
    from fincflow_b200.layers.conv import PaddedConv2d

    torch.manual_seed(3)
    unit = FastFlowUnit(12, 12, (3, 3)).cuda()
    x = torch.randn(6, 12, 8, 8, device="cuda", requires_grad=True)
    z, logdet = unit(x)
    assert logdet == 0.0 and isinstance(logdet, float)     # fastflow.py:34-50
    dz = torch.randn_like(z)
    z.backward(dz)
    wn = unit.weight.detach().cpu().numpy()
    assert rel_err(x.grad.cpu().numpy(), fo.backward_input(dz.cpu().numpy(), wn)) <= REL_TOL
    raw = fo.backward_weight(dz.cpu().numpy(), x.detach().cpu().numpy(), (3, 3), apply_mask=False)
    assert rel_err(unit.weight.grad.cpu().numpy(), raw) <= REL_TOL   # autograd gives the RAW gradient
    torch.nn.Sequential(unit).apply(clear_grad)                     # train/experiment.py:250
    assert rel_err(unit.weight.grad.cpu().numpy(), fo.apply_grad_mask(raw)) <= REL_TOL
    xr = unit.reverse(z.detach())
    assert isinstance(xr, torch.Tensor) and (xr - x.detach()).abs().max().item() <= RT_TOL
    # fused-mask + tensor-logdet mode
    unit2 = FastFlowUnit(12, 12, (3, 3), mask_in_backward=True, logdet_mode="tensor").cuda()
    unit2.load_state_dict(unit.state_dict())
    x2 = x.detach().clone().requires_grad_(True)
    z2, ld2 = unit2(x2)
    assert ld2.shape == (6,) and ld2.abs().max().item() == 0.0
    z2.backward(dz)
    assert torch.equal(z2, z) and torch.equal(unit2.weight.grad, unit.weight.grad)
    # quadrant views behave like the reference's sub-modules
    zq, _ = unit.conv_bl(x.detach()[:, 6:9].contiguous())
    assert torch.equal(zq, z.detach()[:, 6:9])

    for order in ("TL", "TR", "BL", "BR"):
        conv = PaddedConv2d(4, 4, (3, 3), order=order).cuda()
        xc = torch.randn(5, 4, 14, 14, device="cuda", requires_grad=True)
        zc, ld = conv(xc)
        assert ld == 0.0
        zc.sum().backward()
        conv.reset_gradients()
        g = conv.conv.weight.grad.cpu().numpy()
        assert np.array_equal(g == 0, conv.mask.numpy() == 0)
        y, zero = conv.reverse(zc.detach())                      # tuple, layers/conv.py:163
        assert zero == 0 and (y - xc.detach()).abs().max().item() <= RT_TOL
        wn = conv.conv.weight.detach().cpu().numpy()
        assert rel_err(zc.detach().cpu().numpy(), fo.forward(xc.detach().cpu().numpy(), wn, (order,))) <= REL_TOL


def test_reference_pybind_call_site_runs_on_the_shim(nat):
    """FastFlowUnit.reverse_level2 of the reference (fastflow/fastflow.py:78-100), restated
    line by line, calling the drop-in for its pybind `cinc_cuda_level2.inverse`."""
    from fincflow_b200 import compat
    from fincflow_b200.fastflow import FastFlowUnit

    torch.manual_seed(7)
    unit = FastFlowUnit(24, 24, (3, 3)).cuda()
    x = torch.randn(5, 24, 8, 8, device="cuda")
    z, _ = unit(x)
    w = [unit.conv_tl.conv.weight, unit.conv_tr.conv.weight, unit.conv_bl.conv.weight, unit.conv_br.conv.weight]
    k_tl, k_tr = w[0].data, torch.flip(w[1].data, [3])
    k_bl, k_br = torch.flip(w[2].data, [2]), torch.flip(w[3].data, [2, 3])
    kernel = torch.cat([k_tl, k_tr, k_bl, k_br], dim=0).contiguous()
    o_tl, o_tr, o_bl, o_br = torch.chunk(z.detach(), 4, dim=1)
    xin = torch.cat([o_tl, torch.flip(o_tr, [3]), torch.flip(o_bl, [2]), torch.flip(o_br, [2, 3])], dim=1).contiguous()
    y = torch.zeros_like(xin)
    y = compat.cinc_cuda_level2.inverse(xin, kernel, y)[0]
    o_tl, o_tr, o_bl, o_br = torch.chunk(y, 4, dim=1)
    y = torch.cat([o_tl, torch.flip(o_tr, [3]), torch.flip(o_bl, [2]), torch.flip(o_br, [2, 3])], dim=1)
    assert (y - x).abs().max().item() <= RT_TOL
    assert rel_err(y.cpu().numpy(), unit.reverse(z.detach()).cpu().numpy()) <= REL_TOL
    # level-1 shim: one TL-form convolution over all channels (layers/conv.py:191-218)
    from fincflow_b200.layers.conv import PaddedConv2d

    conv = PaddedConv2d(4, 4, (3, 3), order="BR").cuda()
    xc = torch.randn(3, 4, 14, 14, device="cuda")
    zc, _ = conv(xc)
    kern = torch.flip(conv.conv.weight.data, [2, 3]).contiguous()
    yc = compat.cinc_cuda_level1.inverse(torch.flip(zc.detach(), [2, 3]).contiguous(), kern, torch.zeros_like(xc))[0]
    assert (torch.flip(yc, [2, 3]) - xc).abs().max().item() <= RT_TOL


@pytest.mark.parametrize("shape", [(256, 12, 16, 16, 3), (256, 24, 8, 8, 3), (256, 48, 4, 4, 3), (37, 12, 32, 32, 3),
                                   (64, 96, 4, 4, 5), (16, 4, 14, 14, 3)],
                         ids=lambda s: "B{}C{}_{}x{}_k{}".format(*s))
def test_prepared_weight_tables_are_bit_identical(nat, shape):
    """finc_prepare_weights_f32 + FINC_FLAG_PREPARED == staging the raw weights in the kernel"""
    B, CT, H, W, k = shape
    torch.manual_seed(5)
    from fincflow_b200.fastflow import FastFlowUnit

    n_units = 3
    ws = torch.stack([FastFlowUnit(CT, CT, (k, k)).weight.detach() for _ in range(n_units)]).cuda()
    ws[1, 0, 0, k - 1, k - 1] = 1.7  # a non-unit diagonal so that logdet is not trivially 0
    x = torch.randn(B, CT, H, W, device="cuda")
    for kind, fn in ((nat.PREP_FORWARD, None), (nat.PREP_BACKWARD_INPUT, nat.backward_input), (nat.PREP_INVERSE, nat.inverse)):
        nb = nat.prepared_weights_bytes(kind, B, 4, CT // 4, H, W, k, k)
        assert nb > 0 and nb % 128 == 0
        tables = torch.empty((n_units, nb), dtype=torch.uint8, device="cuda")
        nat.prepare_weights(ws, tables, kind, B, H, W)
        for u in range(n_units):
            if kind == nat.PREP_FORWARD:
                z0, ld0 = nat.forward(x, ws[u])
                z1, ld1 = nat.forward(x, None, prepared=tables[u], ksize=(k, k))
                assert torch.equal(z0, z1)
                assert torch.allclose(ld0, ld1, rtol=1e-6, atol=1e-6)
                if u == 1:
                    assert abs(ld1[0].item() - H * W * np.log(1.7)) <= 1e-3 * H * W
            else:
                assert torch.equal(fn(x, ws[u]), fn(x, None, prepared=tables[u], ksize=(k, k)))
    # shapes outside the tiled kernels report 0 bytes instead of a table
    assert nat.prepared_weights_bytes(nat.PREP_FORWARD, 4, 1, 5, 6, 6, 2, 3) == 0


def _fuzz_cases(n, seed=20261018):
    """random shapes over every dispatch family: channel counts in and out of the instantiated sets,
    square and non-square images and kernels, 1..4 groups with random corner orders, ragged batches"""
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n):
        C = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 12, 24]))
        G = int(rng.choice([1, 2, 3, 4]))
        H = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 12, 14, 16, 20]))
        W = int(rng.choice([1, 2, 3, 4, 6, 7, 8, 12, 14, 16, 32, 36]))
        k = (3, 3) if rng.random() < 0.6 else [(5, 5), (2, 2), (1, 3), (3, 2), (4, 4)][int(rng.integers(5))]
        if C * H * W > 4096:
            H = max(1, 4096 // (C * W))
        B = int(rng.choice([1, 2, 3, 5, 17, 64, 130, 600])) if C * H * W < 800 else int(rng.choice([1, 2, 3, 9, 33]))
        orders = tuple(int(o) for o in rng.integers(0, 4, size=G))
        cases.append((B, G, C, H, W, k[0], k[1], orders))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(48), ids=lambda c: "B{}G{}C{}_{}x{}_k{}x{}_o{}".format(*c[:7], "".join(map(str, c[7]))))
def test_fuzz_against_oracle(nat, case):
    """every kernel family on random shapes vs the oracle (same tolerances as the named cases)"""
    B, G, C, H, W, kH, kW, orders = case
    rng = np.random.default_rng(abs(hash(case[:7])) % (2**32))
    w = np.concatenate([fo.init_weight(C, (kH, kW), o, rng) for o in orders], 0)
    x = rng.normal(size=(B, G * C, H, W)).astype(np.float32)
    dz = rng.normal(size=x.shape).astype(np.float32)
    zs = rng.normal(size=x.shape).astype(np.float32)
    r = run_all(nat, x, w, dz, zs, orders)
    assert rel_err(r["z"], fo.forward(x, w, orders)) <= REL_TOL
    assert rel_err(r["dx"], fo.backward_input(dz, w, orders)) <= REL_TOL
    assert rel_err(r["dw"], fo.backward_weight(dz, x, (kH, kW), orders)) <= REL_TOL
    assert rel_err(r["x_zs"], fo.inverse(zs, w, orders)) <= REL_TOL
    assert np.abs(r["x_rt"] - x).max() <= RT_TOL


@pytest.mark.parametrize("shape", [(37, 12, 16, 16, 3), (1031, 24, 8, 8, 3), (300, 48, 4, 4, 3), (9, 96, 4, 4, 5), (33, 8, 7, 7, 3),
                                   (5, 12, 32, 32, 5), (2100, 12, 16, 16, 3), (3, 20, 6, 36, 3)],
                         ids=lambda s: "B{}C{}_{}x{}_k{}".format(*s))
def test_no_out_of_bounds_writes(nat, shape):
    """every output lives inside a larger buffer filled with a sentinel: the kernels (TMA bulk stores,
    vector stores, ragged last items) must leave the guard zones on both sides untouched"""
    B, CT, H, W, k = shape
    torch.manual_seed(B + CT)
    from fincflow_b200.fastflow import FastFlowUnit

    GUARD, SENT = 4096, 12345.678

    def guarded(n):
        buf = torch.full((n + 2 * GUARD,), SENT, device="cuda")
        return buf, buf[GUARD:GUARD + n]

    def check(buf, n, what):
        assert bool((buf[:GUARD] == SENT).all()) and bool((buf[GUARD + n:] == SENT).all()), what

    w = FastFlowUnit(CT, CT, (k, k)).weight.detach().cuda()
    x = torch.randn(B, CT, H, W, device="cuda")
    dz = torch.randn_like(x)
    n = x.numel()
    for name, fn in (("forward", lambda o: nat.forward(x, w, out=o, want_logdet=False)),
                     ("backward_input", lambda o: nat.backward_input(dz, w, out=o)),
                     ("inverse", lambda o: nat.inverse(x, w, out=o)),
                     ("inverse_wave", lambda o: nat.inverse(x, w, out=o, flags=nat.FLAG_WAVE_SMEM)),
                     ("inverse_generic", lambda o: nat.inverse(x, w, out=o, flags=nat.FLAG_GENERIC_TILED))):
        buf, view = guarded(n)
        fn(view.view_as(x))
        torch.cuda.synchronize()
        check(buf, n, name)
        assert not bool((view == SENT).any()), name + ": unwritten output"
    buf, view = guarded(w.numel())
    nat.backward_weight(dz, x, (k, k), out=view.view_as(w))
    torch.cuda.synchronize()
    check(buf, w.numel(), "backward_weight")
    A = torch.linalg.qr(torch.randn(CT, CT, device="cuda"))[0].contiguous()
    buf, view = guarded(n)
    nat.affine1x1(x, A, torch.randn(CT, device="cuda"), out=view.view_as(x))
    torch.cuda.synchronize()
    check(buf, n, "affine1x1")
    if H % 2 == 0 and W % 2 == 0:
        buf, view = guarded(n)
        nat.squeeze(x, out=view.view(B, 4 * CT, H // 2, W // 2))
        torch.cuda.synchronize()
        check(buf, n, "squeeze")


@pytest.mark.parametrize("case", [(300, 48, 8, 8, 5), (300, 48, 8, 8, 3), (257, 96, 4, 4, 5), (64, 96, 4, 4, 3), (100, 48, 4, 4, 3),
                                  (33, 24, 8, 8, 3), (5, 16, 4, 4, 3)])
def test_dense_inverse_matches_wavefront_inverse(case):
    """finc_inverse_dense_f32 (x = L^-1 z as tensor-core GEMMs, L^-1 from the wavefront kernel on the identity)
    is the same map as finc_inverse_f32: <= 1e-5 relative, <= 1e-4 round trip, and agrees with the oracle"""
    from fincflow_b200 import _native
    from fincflow_b200.fastflow import FastFlowUnit

    B, C, H, W, k = case
    torch.manual_seed(B + C)
    unit = FastFlowUnit(C, C, (k, k)).cuda()
    w = unit.weight.detach()
    assert _native.inverse_dense_bytes(4, C // 4, H, W) > 0
    blob = _native.inverse_dense_prepare(w, H, W)
    z = torch.randn(B, C, H, W, device="cuda")
    xd = _native.inverse_dense(z, blob)
    xw = _native.inverse(z, w)
    assert rel_err(xd.cpu().numpy(), xw.cpu().numpy()) <= REL_TOL
    assert elementwise_close(xd.cpu().numpy(), xw.cpu().numpy())
    zz, _ = _native.forward(xd, w, want_logdet=False)
    assert float((zz - z).abs().max()) <= RT_TOL
    if B <= 64:
        assert rel_err(xd.cpu().numpy(), fo.inverse(z.cpu().numpy(), w.cpu().numpy())) <= REL_TOL
    assert _native.inverse_dense_bytes(4, 3, 16, 16) > 0 and _native.inverse_dense_bytes(4, 3, 32, 32) == 0


def test_fastflowunit_dense_reverse_is_cached_per_weight_version():
    from fincflow_b200.fastflow import FastFlowUnit

    torch.manual_seed(1)
    unit = FastFlowUnit(48, 48, (5, 5)).cuda()
    z = torch.randn(130, 48, 4, 4, device="cuda")
    with torch.no_grad():
        want = unit.reverse(z)
        unit.dense_reverse = True
        got = unit.reverse(z)
        blob = unit._dense_blob
        assert unit.reverse(z) is not None and unit._dense_blob is blob          # cached
        assert rel_err(got.cpu().numpy(), want.cpu().numpy()) <= REL_TOL
        unit.weight.mul_(1.01)                                                    # new version -> new table
        got2 = unit.reverse(z)
        unit.dense_reverse = False
        assert rel_err(got2.cpu().numpy(), unit.reverse(z).cpu().numpy()) <= REL_TOL
        assert rel_err(got2.cpu().numpy(), want.cpu().numpy()) > 1e-4
