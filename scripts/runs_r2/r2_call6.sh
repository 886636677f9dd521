#!/bin/bash
# Round 2, GPU call 6: two MMA-issuing warps (step + linear kernels), whole-K partial sums (NFK_STEP_DRAIN=8), AR transform on the
# step kernel (growing degree-sorted sub-networks).
mkdir -p gpurun_out
LOG=gpurun_out/call6.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
for s in t1 t3 t4 c0 c1 c2 c3 c4 c6 c8; do run python scripts/step_check.py $s; done
NFK_STEP_DRAIN=8 run python scripts/step_check.py c3
NFK_STEP_DRAIN=8 run python scripts/step_check.py c2
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 >> $LOG
run python scripts/linear_only.py 784 784 pair
NFK_LINEAR_MMA_WARPS=1 run python scripts/linear_only.py 784 784 pair
run python scripts/linear_only.py 256 256 pair
bench() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spline-roofline $EXTRA 2>gpurun_out/bench6_$tag.err | tail -1 > gpurun_out/bench6_$tag.json
  python - "$tag" >> $LOG <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench6_%s.json" % sys.argv[1]))
    print(sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"], "parity", d["parity_check"]["rel_err"], d["parity_check"]["ok"], "e2e", d["e2e"]["value"], "extra", d.get("extra"))
except Exception as e:
    print("bench failed", sys.argv[1], e)
PY
  tail -3 gpurun_out/bench6_$tag.err >> $LOG
}
EXTRA="" bench default A=1
EXTRA="--no-extras" bench mma1 NFK_STEP_MMA_WARPS=1 NFK_LINEAR_MMA_WARPS=1
EXTRA="--no-extras" bench drain8 NFK_STEP_DRAIN=8
EXTRA="--no-extras" bench ewg2 NFK_STEP_EWG=2
echo "=== ncu" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_coupling_step -s 4 -c 1 -o gpurun_out/ncu_step_r2e -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_f16x3 -s 4 -c 1 -o gpurun_out/ncu_linear_r2e -f python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
cat $LOG
