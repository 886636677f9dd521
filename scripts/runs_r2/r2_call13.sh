#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/call13.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 >> gpurun_out/call13.log
cat gpurun_out/call13.log
bash scripts/r2_profile.sh
