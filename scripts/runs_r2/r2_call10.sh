#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call10.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
NFK_STEP_DEBUG_X=1 run python scripts/step_check.py c3
NFK_STEP_DEBUG_X=1 run python scripts/step_check.py c3
NFK_STEP_DEBUG_X=2 run python scripts/step_check.py c3
NFK_STEP_DEBUG_X=1 NFK_STEP_DRAIN=4 run python scripts/step_check.py c3
cat $LOG
