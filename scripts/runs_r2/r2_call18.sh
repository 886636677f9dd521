#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call18.log
: > $LOG
timeout 600 python -m pytest tests -q -m gpu -k "sharding or stream or e2e" 2>&1 | tail -4 >> $LOG
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check"
run() { echo "--- $*" >> $LOG; env "$@" timeout 300 $B ${EXTRA} 2>> $LOG | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'])
" >> $LOG 2>&1; }
EXTRA="" run A=1
EXTRA="" run NFLOWS_B200_STREAM_RAMP=0
EXTRA="--e2e-chunk 262144" run A=1
EXTRA="--e2e-chunk 262144" run NFLOWS_B200_STREAM_RAMP=0
EXTRA="--e2e-chunk 65536" run A=1
EXTRA="--e2e-chunk 524288" run A=1
cat $LOG
