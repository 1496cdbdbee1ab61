"""time the tensor-core 1x1 conv (plain GEMM) at the coupling shapes (run under gpurun)
   python tools/gemm_time.py [B H W Cin N taps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native

a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [256, 16, 16, 512, 512, 1]
B, H, W, Cin, N, taps = a
dev = torch.device("cuda:0")
k = 3 if taps == 9 else 1
x = torch.randn(B, H, W, Cin, device=dev)
w = torch.randn(N, Cin, k, k, device=dev) / (Cin * taps) ** 0.5
wp = _native.tc_conv_prepare_weights(w, 0)
bias = torch.zeros(N, device=dev)
for flags, name in ((0, "3xTF32"), (_native.FLAG_TF32_1PASS, "1xTF32")):
    y = _native.tc_conv_nhwc(x, wp, bias, N, taps, relu=True, flags=flags)
    for _ in range(3):
        _native.tc_conv_nhwc(x, wp, bias, N, taps, relu=True, flags=flags, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _native.tc_conv_nhwc(x, wp, bias, N, taps, relu=True, flags=flags, out=y)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * B * H * W * Cin * taps * N
    print(f"dbg={os.environ.get('FINC_TC_DBG', '0')} {name} [{B},{H},{W}] {Cin}x{taps}->{N}: {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s fp32-equivalent")
