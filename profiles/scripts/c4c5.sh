#!/bin/bash
# usage: c4c5.sh N -- the 1.3 M-cell configuration and the kNN/SNN sweep at N ranks
N=$1
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"; fi
timeout 900 $L bench.py --gpus $N --config c4 --steps 3 --warmup 3 --parity off --no-cpu-baseline > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
echo "c4 rc=$?"; grep '^{' gpurun_out/r2_bench_c4_n$N.json | head -c 600; tail -3 gpurun_out/r2_bench_c4_n$N.err
timeout 600 $L bench.py --gpus $N --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
echo "c5 rc=$?"; grep '^{' gpurun_out/r2_bench_c5_n$N.json | head -c 900; tail -3 gpurun_out/r2_bench_c5_n$N.err
