"""Run only the HBM-bound spline kernel (nfk_rqs_rows) at the bench shape -- target for `ncu -k regex:rqs_rows`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(bench.spline_hbm_roofline(torch.device("cuda:0"), bench.load_peaks(), iters=3))
