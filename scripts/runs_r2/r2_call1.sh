#!/bin/bash
# Round 2, GPU call 1: first hardware run of the round-1 persistent trunk kernel (resident swizzled operand written by the
# epilogue), A/B bench, and dedicated ncu captures of the 784x784 affine GEMM and the HBM-bound spline kernel.
mkdir -p gpurun_out
{
echo "=== trunk kernel tests"
NFLOWS_B200_TRUNK_KERNEL=1 timeout 300 python -m pytest tests/test_trunk_kernel.py -m gpu -q -x 2>&1 | tail -15
for v in 1 0; do
  NFLOWS_B200_TRUNK_KERNEL=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spline-roofline 2>/dev/null | tail -1 > gpurun_out/bench_trunk$v.json
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_trunk%s.json" % sys.argv[1]))
    print("trunk_kernel", sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"])
except Exception as e:
    print("bench failed", sys.argv[1], e)
PY
done
echo "=== linear 784x784 timing"
timeout 120 python scripts/linear_only.py 784 784 pair
timeout 120 python scripts/linear_only.py 784 784 y
echo "=== ncu linear 784"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linear_f16x3 -s 3 -c 1 -o gpurun_out/ncu_linear784_r2 -f python scripts/linear_only.py 784 784 pair 2>&1 | tail -3
echo "=== ncu rqs_rows"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rqs_rows -s 3 -c 1 -o gpurun_out/ncu_rqs_rows_r2 -f python scripts/spline_only.py 2>&1 | tail -3
} > gpurun_out/call1.log 2>&1
tail -40 gpurun_out/call1.log
