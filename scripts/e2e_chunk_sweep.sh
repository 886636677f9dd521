#!/bin/bash
# end-to-end leg of bench.py at several host->device chunk sizes
for c in "$@"; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spline-roofline --e2e-chunk $c 2>/dev/null | tail -1 > /tmp/_e2e.json
  python - "$c" <<'PY'
import json, sys
d = json.load(open("/tmp/_e2e.json"))
print("e2e_chunk", sys.argv[1], "device %.0f" % d["value"], "e2e %.0f samples/s" % d["e2e"]["value"], "e2e ms/step %.1f" % d["e2e"]["ms_per_step"])
PY
done
