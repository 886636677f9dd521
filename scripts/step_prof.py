"""Where the cycles of rq_coupling_step_kernel go (per-CTA counters written through NFK_STEP_PROF): the MMA-issuing thread's waits
by phase, and epilogue warp 5's time by activity.  cfg-3 layer shape: D = 784, H = 256, 2 blocks, K = 8."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nflows_b200 import config
from nflows_b200 import kernels as K
from nflows_b200.flows import recipes

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
flow = recipes.perturb_(recipes.rq_nsf(784, 256, num_layers=1)).eval().to(dev)
coupling = flow._transform._transforms[2]
x = torch.randn(rows, 784, device=dev)
with torch.no_grad():
    for _ in range(2):
        coupling(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        coupling(x)
    e1.record()
    torch.cuda.synchronize()
    print("rows %d: %.3f ms per coupling call (gather + step kernel + scatter)" % (rows, e0.elapsed_time(e1) / 3))
    prof = torch.zeros(148 * 17, dtype=torch.int64, device=dev)
    os.environ["NFK_STEP_PROF"] = hex(prof.data_ptr())
    coupling(x)
    torch.cuda.synchronize()
    os.environ.pop("NFK_STEP_PROF")
p = prof[:148 * 16].view(148, 16).double().cpu()
store = prof[148 * 16:].double().cpu().mean()
m = p.mean(dim=0)
names = ["MMA thread total", "  trunk: wait drained accumulator", "  trunk: wait operands", "  trunk: wait activation (layer dependency)",
         "  final: wait drained accumulator", "  final: wait operands", "  final: wait activation",
         "epilogue warp total", "  trunk: prologue (skip + bias)", "  trunk: wait partial sum", "  trunk: drains", "  trunk: layer epilogue (relu, pair, R, skip)",
         "  final: wait partial sum", "  final: drain", "  final: input tile + bias", "  final: spline"]
for i, nm in enumerate(names):
    base = m[0] if i < 7 else m[7]
    print("  %-48s %10.0f cycles  %5.1f %%" % (nm, m[i], 100 * m[i] / base))
print("  %-48s %10.0f cycles  %5.1f %%" % ("  final: staging, barrier, store", store, 100 * store / m[7]))
