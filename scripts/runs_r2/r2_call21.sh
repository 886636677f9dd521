#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call21.log
: > $LOG
timeout 300 python scripts/linear_prof.py 262144 >> $LOG 2>&1
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 >> $LOG
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], d['timeline_ms_per_step'], d['parity_check']['rel_err'])
" >> $LOG 2>&1
cat $LOG
