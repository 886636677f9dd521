#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call14.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 >> $LOG
echo "=== bench" >> $LOG
timeout 1200 python bench.py > gpurun_out/bench_call14.json 2>> $LOG
python - >> $LOG <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call14.json").read().strip().splitlines()[-1])
print("samples/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "traffic", d["roofline"]["traffic"])
for k, v in d.get("extra", {}).items():
    print(k, v)
PY
cat $LOG | cut -c1-700
