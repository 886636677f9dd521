#!/bin/bash
# validates the pair-only output mode of the fused coupling kernel: GPU parity tests and bench, with and without it
NFLOWS_B200_PAIR_ONLY=1 timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -2
for v in 1 0; do
  NFLOWS_B200_PAIR_ONLY=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spline-roofline 2>/dev/null | tail -1 > /tmp/_po.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/_po.json"))
print("pair_only", sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"])
PY
done
