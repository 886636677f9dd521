#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call31.log
: > $LOG
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check"
run() { echo "--- $*" >> $LOG; env "$@" timeout 300 $B 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], d['timeline_ms_per_step'])
" >> $LOG 2>&1; }
run A=1
run NFK_STEP_X_PREFETCH=1
run A=1
cat $LOG
