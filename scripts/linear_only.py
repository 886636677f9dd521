"""One trunk-shaped tensor-core dense layer (n=2^18, fp16 split-pair operands, residual, fp32 out) -- ncu
target and quick timing.  usage: linear_only.py [K] [N] [mode]   mode: y (fp32 output only, default) | pair (fp32 + fp16 pair output)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nflows_b200 import kernels as K
dev = torch.device("cuda:0")
n = 1 << 18
k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
o = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "y"
x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) / k ** 0.5; b = torch.randn(o, device=dev)
r = torch.randn(n, o, device=dev)
wp = K.split_f16(w, K.weight_exp(w))
y = torch.empty(n, o, device=dev)
xp = K.split_f16(x, 6, relu=True)
pair = K.Pair16.empty(n, o, 6, dev)
run = lambda: K.linear_f16x3(xp, wp, b, residual=r, want_y=True, y_out=y, want_split=mode == "pair", pair_out=pair if mode == "pair" else None)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("linear %dx%d n=%d %s: %.3f ms  %.1f TFLOP/s (algorithmic)" % (k, o, n, mode, ms, 2 * n * k * o / ms / 1e9))
