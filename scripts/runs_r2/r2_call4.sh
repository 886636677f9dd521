#!/bin/bash
# Round 2, GPU call 4: lean spline (binary bin search) in every spline kernel, two-ring linear kernel, y_first_col, auto activation
# exponent; correctness first, then timings and profiles.
mkdir -p gpurun_out
LOG=gpurun_out/call4.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
for s in c0 c1 c2 c3 c4 c6 c8; do run python scripts/step_check.py $s; done
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 >> $LOG
echo "=== linear 784x784" >> $LOG
run python scripts/linear_only.py 784 784 pair
NFK_LINEAR_NINNER=0 run python scripts/linear_only.py 784 784 pair
run python scripts/linear_only.py 256 256 pair
echo "=== spline hbm" >> $LOG
run python scripts/spline_only.py
echo "=== bench" >> $LOG
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench4.err | tail -1 > gpurun_out/bench4.json
python - >> $LOG <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench4.json"))
    print("bench samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"], d["timeline_ms_per_step"], "launches", d["gpu_launches"])
    print("parity", d.get("parity_check")); print("e2e", d["e2e"]); print("roofline", d.get("roofline")); print("spline", d.get("roofline_spline"))
    print("extra", d.get("extra")); print("torch_cuda", d.get("torch_cuda_baseline")); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench failed", e)
PY
tail -5 gpurun_out/bench4.err >> $LOG
echo "=== ncu" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_coupling_step -s 4 -c 1 -o gpurun_out/ncu_step_r2c -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_f16x3 -s 4 -c 1 -o gpurun_out/ncu_linear_r2c -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rqs_rows -s 3 -c 1 -o gpurun_out/ncu_rqs_rows_r2c -f python scripts/spline_only.py 2>&1 | tail -2 >> $LOG
cat $LOG
