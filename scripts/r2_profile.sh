#!/bin/bash
# Evidence pass on ONE GPU: launch list of the bench command, then one `ncu --set full` capture per hot kernel.
mkdir -p gpurun_out
LOG=gpurun_out/profile.log
: > $LOG
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check"
echo "=== launch list" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r2.csv $B > gpurun_out/bench_under_ncu.log 2>&1
echo "rc=$?" >> $LOG
S="python bench.py --steps 1 --warmup 1 --rows 262144 --no-cpu-baseline --no-spline-roofline --no-extras --no-parity-check"
for k in rq_coupling_step linear_f16x3; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/ncu_${k}_r2final -f $S 2>&1 | tail -2 >> $LOG
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rqs_rows -s 2 -c 1 -o gpurun_out/ncu_rqs_rows_r2final -f python bench.py --steps 1 --warmup 1 --rows 131072 --no-cpu-baseline --no-extras --no-parity-check 2>&1 | tail -2 >> $LOG
echo "=== clean bench line (no profiler)" >> $LOG
timeout 900 python bench.py > gpurun_out/bench_final.json 2>> $LOG
tail -c 600 gpurun_out/bench_final.json >> $LOG
cat $LOG | cut -c1-600
