"""SM clock / power while the 3xTF32 GEMM runs back to back (run under gpurun)"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fincflow_b200 import _native

dev = torch.device("cuda:0")
taps = int(sys.argv[1]) if len(sys.argv) > 1 else 9
flags = _native.FLAG_TF32_1PASS if len(sys.argv) > 2 and sys.argv[2] == "1" else 0
B, H, W, Cin, N = 256, 16, 16, 512, 512
k = 3 if taps == 9 else 1
x = torch.randn(B, H, W, Cin, device=dev)
w = torch.randn(N, Cin, k, k, device=dev) / (Cin * taps) ** 0.5
wp = _native.tc_conv_prepare_weights(w, 0)
bias = torch.zeros(N, device=dev)
y = _native.tc_conv_nhwc(x, wp, bias, N, taps, relu=True, flags=flags)
samples = []
stop = False
def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active",
                              "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.2)
th = threading.Thread(target=sampler); th.start()
t0 = time.time()
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
while time.time() - t0 < 4.0:
    e0.record()
    for _ in range(50):
        _native.tc_conv_nhwc(x, wp, bias, N, taps, relu=True, flags=flags, out=y)
    e1.record()
    torch.cuda.synchronize()
    n += 1
    last = e0.elapsed_time(e1) / 50 * 1e3
stop = True; th.join()
print("last us per launch", round(last, 1))
for s in samples[::2]:
    print(s)
