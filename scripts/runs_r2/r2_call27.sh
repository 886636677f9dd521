#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call27.log
: > $LOG
timeout 300 python scripts/step_prof.py 262144 >> $LOG 2>&1
NFK_CLUSTER=1 timeout 300 python scripts/step_prof.py 262144 >> $LOG 2>&1
for i in 1 2 3; do timeout 120 python scripts/step_check.py c3 2>&1 | head -3 | cut -c1-200 >> $LOG; done
cat $LOG
