#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call11.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run python scripts/step_check.py c3
NFK_STEP_DUMP=1 run python scripts/step_check.py c3
NFK_STEP_DUMP=1 run python scripts/step_check.py c3
NFK_STEP_DUMP=1 NFK_CLUSTER=1 run python scripts/step_check.py c3
cat $LOG
