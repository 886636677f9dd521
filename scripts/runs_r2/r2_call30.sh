#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call30.log
: > $LOG
for s in c0 c1 c2 c3 c3 c4 c5 c6 c7 c8; do timeout 120 python scripts/step_check.py $s 2>&1 | head -4 | cut -c1-220 >> $LOG; done
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras"
run() { echo "--- $*" >> $LOG; env "$@" timeout 300 $B 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], d['timeline_ms_per_step'], d['parity_check']['rel_err'])
" >> $LOG 2>&1; }
run A=1
run NFK_STEP_K16=0
run A=1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $LOG
cat $LOG
