#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/final.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> $LOG
echo "=== bench" >> $LOG
timeout 1200 python bench.py > gpurun_out/bench_final.json 2>> $LOG
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_final_reference.json 2>> $LOG
python - >> $LOG <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("samples/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity_check"]["rel_err"], d["timeline_ms_per_step"], d["clocks"])
r = json.loads(open("gpurun_out/bench_final_reference.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"])
PY
cat $LOG | cut -c1-500
