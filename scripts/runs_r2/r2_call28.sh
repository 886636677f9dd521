#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call28.log
: > $LOG
for s in c0 c1 c2 c3 c3 c4 c5 c6 c7 c8 t4; do timeout 120 python scripts/step_check.py $s 2>&1 | head -4 | cut -c1-220 >> $LOG; done
timeout 300 python scripts/step_prof.py 262144 >> $LOG 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $LOG
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], d['timeline_ms_per_step'], d['parity_check']['rel_err'])
" >> $LOG 2>&1
cat $LOG
