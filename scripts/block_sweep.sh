#!/bin/bash
# bench.py at several rows-per-launch-round settings (config.{trunk,affine,coupling}_block_rows)
for br in "$@"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spline-roofline --block-rows $br 2>/dev/null | tail -1 > /tmp/_sweep.json
  python - "$br" <<'PY'
import json, sys
d = json.load(open("/tmp/_sweep.json"))
print("block_rows", sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"])
PY
done
