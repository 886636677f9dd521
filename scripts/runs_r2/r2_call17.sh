#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call17.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 >> $LOG
echo "=== bench" >> $LOG
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_call17.json 2>> $LOG
python - >> $LOG <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call17.json").read().strip().splitlines()[-1])
print("samples/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity_check"]["rel_err"])
PY
cat $LOG | cut -c1-400
