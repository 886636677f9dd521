#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call12.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 300 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run python scripts/step_check.py c3
run python scripts/step_check.py c3
NFK_STEP_DUMP=1 run python scripts/step_check.py c3
NFK_STEP_DUMP=1 NFK_CLUSTER=1 run python scripts/step_check.py c3
for s in c0 c1 c2 c4 c5 c6 c7 c8 t4; do run python scripts/step_check.py $s; done
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 >> $LOG
NFLOWS_B200_STEP_KERNEL=0 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py > gpurun_out/bench_call12.json 2>> $LOG
python - >> $LOG <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call12.json").read().strip().splitlines()[-1])
print("samples/s", d["value"], "ms/step", d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d.get("timeline_ms_per_step"), "parity", d.get("parity_check"), "e2e", d["e2e"]["value"])
print("roofline", d["roofline"])
for k in ("extra_workloads", "torch_cuda_baseline", "cpu_baseline"):
    print(k, d.get(k))
PY
cat $LOG | cut -c1-400
