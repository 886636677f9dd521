"""Where the cycles of linear_f16x3_kernel go (per-CTA counters written through NFK_LINEAR_PROF): MMA thread waiting for drained
accumulators / for operands, epilogue warp 4 waiting for partial sums / draining / writing the tile out."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nflows_b200 import dense as D
from nflows_b200 import kernels as K

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
k = n = 784
torch.manual_seed(0)
x = torch.randn(rows, k, device=dev)
w = torch.randn(n, k, device=dev) / 28
b = torch.randn(n, device=dev)
xp = K.split_f16(x, D.act_exp())
wp = K.split_f16(w, K.weight_exp(w))
y = torch.empty(rows, n, device=dev)
yp = K.Pair16.empty(rows, n, D.act_exp(), dev)


def launch():
    # the affine fold's call shape: fp32 for the transformed block, the pair for the identity block
    K.linear_f16x3(xp, wp, b, want_y=True, y_out=y, want_split=True, split_cols=392, pair_out=yp, y_first_col=392)


for _ in range(3):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    launch()
e1.record()
torch.cuda.synchronize()
print("rows %d: %.3f ms per launch" % (rows, e0.elapsed_time(e1) / 5))
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
os.environ["NFK_LINEAR_PROF"] = hex(prof.data_ptr())
launch()
torch.cuda.synchronize()
os.environ.pop("NFK_LINEAR_PROF")
p = prof.view(148, 16).double().cpu()
names = ["mma total", "mma wait tempty", "mma wait operands", "epi total", "epi wait tfull", "epi drain", "epi tile output", "epi tile prologue",
         "  prologue: bias", "  output: store waits", "  output: fp32 chunks", "  output: pair chunks"]
m = p.mean(dim=0)
for i, nm in enumerate(names):
    base = m[0] if i < 3 else m[3]
    print("  %-20s %10.0f cycles  %5.1f %%" % (nm, m[i], 100 * m[i] / base))
