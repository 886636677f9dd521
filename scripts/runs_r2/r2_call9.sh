#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/call9.log
: > $LOG
run() { echo "--- $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run python scripts/step_check.py c3
run python scripts/step_check.py c3
NFK_STEP_DRAIN=4 run python scripts/step_check.py c3
NFK_STEP_TRUNK_DRAIN=2 run python scripts/step_check.py c3
NFK_STEP_NO_TMA_X=1 run python scripts/step_check.py c3
NFK_CLUSTER=1 run python scripts/step_check.py c3
NFK_STEP_EWG=2 run python scripts/step_check.py c3
cat $LOG
