"""Per-kernel device-time breakdown of one Flow.log_prob step (torch.profiler/CUPTI; not a bench number)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device("cuda:0")
flow = bench.build_flow().to(dev)
x = torch.randn(rows, bench.FEATURES, device=dev)
with torch.no_grad():
    flow.log_prob(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        flow.log_prob(x)
        torch.cuda.synchronize()
print("rows", rows, "backend", os.environ.get("NFLOWS_B200_GEMM", "tc"))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20, max_name_column_width=60))
