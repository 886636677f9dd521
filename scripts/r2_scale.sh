#!/bin/bash
# strong-scaling lines: bash scripts/r2_scale.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/scale_$N.log
: > $LOG
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv >> $LOG 2>&1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2>> $LOG
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2>> $LOG
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> $LOG
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --weak --no-parity-check > gpurun_out/bench_weak_n$N.json 2>> $LOG
fi
tail -5 $LOG
for f in gpurun_out/bench_n$N.json gpurun_out/bench_ref_n$N.json gpurun_out/bench_weak_n$N.json; do [ -f $f ] && tail -1 $f | cut -c1-1500; done
