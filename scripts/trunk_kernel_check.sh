#!/bin/bash
# First job of round 2: validate the experimental persistent trunk kernel (nfk_residual_trunk_f16x3) and measure it.
#   gpurun --timeout 900 -- 'scripts/trunk_kernel_check.sh'
NFLOWS_B200_TRUNK_KERNEL=1 timeout 300 python -m pytest tests/test_trunk_kernel.py -m gpu -q -x 2>&1 | tail -4
NFLOWS_B200_TRUNK_KERNEL=1 timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -2
for v in 1 0; do
  NFLOWS_B200_TRUNK_KERNEL=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spline-roofline 2>/dev/null | tail -1 > /tmp/_tk.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/_tk.json"))
print("trunk_kernel", sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"])
PY
done
