"""One trunk-shaped tensor-core dense layer (n=2^18, 256->256, residual + split outputs) -- ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nflows_b200 import kernels as K
dev = torch.device("cuda:0")
n, k, o = 1 << 18, int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 256
x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) / k ** 0.5; b = torch.randn(o, device=dev)
r = torch.randn(n, o, device=dev)
xp, wp = K.split_tf32(x), K.split_tf32(w)
for _ in range(3):
    K.linear_tf32x3(xp, wp, b, residual=r, want_y=True, want_split=True, split_relu=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    K.linear_tf32x3(xp, wp, b, residual=r, want_y=True, want_split=True, split_relu=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("linear %dx%d n=%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (k, o, n, ms, 2 * n * k * o / ms / 1e9))
