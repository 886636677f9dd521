#!/bin/bash
# Round 2, GPU call 5: TMA-staged inputs + packed staging stores in the step kernel; CTA-pair (cta_group::2) mode of the linear kernel.
mkdir -p gpurun_out
LOG=gpurun_out/call5.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
for s in c0 c1 c2 c3 c4 c7; do run python scripts/step_check.py $s; done
NFK_STEP_NO_TMA_X=1 run python scripts/step_check.py c3
echo "=== pytest gpu (default modes)" >> $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 >> $LOG
echo "=== pair mode: gemm tests + timings" >> $LOG
NFK_CLUSTER=3 timeout 300 python -m pytest tests/test_native_gemm.py -m gpu -q 2>&1 | tail -8 >> $LOG
run python scripts/linear_only.py 784 784 pair
NFK_CLUSTER=3 run python scripts/linear_only.py 784 784 pair
NFK_CLUSTER=3 run python scripts/linear_only.py 256 256 pair
for c in 2 3; do
  NFK_CLUSTER=$c timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spline-roofline --no-extras 2>gpurun_out/bench5_$c.err | tail -1 > gpurun_out/bench5_$c.json
  python - "$c" >> $LOG <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench5_%s.json" % sys.argv[1]))
    print("NFK_CLUSTER", sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"], "parity", d["parity_check"]["rel_err"], d["parity_check"]["ok"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", sys.argv[1], e)
PY
  tail -3 gpurun_out/bench5_$c.err >> $LOG
done
NFK_CLUSTER=3 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 >> $LOG
echo "=== ncu" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_coupling_step -s 4 -c 1 -o gpurun_out/ncu_step_r2d -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
NFK_CLUSTER=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_f16x3 -s 4 -c 1 -o gpurun_out/ncu_linear_r2d -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
cat $LOG
