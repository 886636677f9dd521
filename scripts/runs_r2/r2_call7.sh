#!/bin/bash
# Round 2, GPU call 7: how much of the kernels' time is the TMEM -> register drain of partial sums?  Sweep the partial-sum length
# (K-slabs accumulated in TMEM before a drain) for the linear kernel, the step kernel's trunk and its final layer: time + accuracy.
mkdir -p gpurun_out
LOG=gpurun_out/call7.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 240 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
for s in c1 c3; do run python scripts/step_check.py $s; done
NFK_STEP_MMA_WARPS=2 NFK_CLUSTER=1 run python scripts/step_check.py c3
NFK_STEP_MMA_WARPS=2 run python scripts/step_check.py c1
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 >> $LOG
for d in 2 4 8 25; do
  echo "=== NFK_LINEAR_DRAIN=$d" >> $LOG
  NFK_LINEAR_DRAIN=$d run python scripts/linear_only.py 784 784 pair
  NFK_LINEAR_DRAIN=$d timeout 200 python scripts/gemm_accuracy.py 2>&1 | grep -A1 "K=784 N=784\|K=256 N=256" | grep "f16x3 exp=6\|n=4096" >> $LOG
done
bench() {
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spline-roofline --no-extras 2>gpurun_out/bench7_$tag.err | tail -1 > gpurun_out/bench7_$tag.json
  python - "$tag" >> $LOG <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench7_%s.json" % sys.argv[1]))
    print(sys.argv[1], "samples/s %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["timeline_ms_per_step"], "parity", d["parity_check"]["rel_err"], d["parity_check"]["ok"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", sys.argv[1], e)
PY
  tail -2 gpurun_out/bench7_$tag.err >> $LOG
}
bench base A=1
bench fdrain8 NFK_STEP_DRAIN=8
bench tdrain4 NFK_STEP_TRUNK_DRAIN=4
bench ldrain4 NFK_LINEAR_DRAIN=4
bench all NFK_STEP_DRAIN=8 NFK_STEP_TRUNK_DRAIN=4 NFK_LINEAR_DRAIN=4
bench all8 NFK_STEP_DRAIN=8 NFK_STEP_TRUNK_DRAIN=8 NFK_LINEAR_DRAIN=8
NFK_STEP_DRAIN=8 NFK_STEP_TRUNK_DRAIN=4 NFK_LINEAR_DRAIN=4 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 >> $LOG
cat $LOG
