#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call16.log
: > $LOG
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras"
run() { echo "--- $*" >> $LOG; env "$@" timeout 300 $B 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], d.get('timeline_ms_per_step'), 'parity', d.get('parity_check',{}).get('rel_err'), 'e2e ms', round(d['e2e']['ms_per_step'],1))
" >> $LOG 2>&1; }
run A=1
run NFK_CLUSTER=3
run NFLOWS_B200_STEP_KERNEL=0
run NFLOWS_B200_STEP_KERNEL=0 NFK_CLUSTER=3
cat $LOG
