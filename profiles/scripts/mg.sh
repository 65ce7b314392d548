#!/bin/bash
# usage: mg.sh N  -- multi-GPU parity check + C3 bench at N ranks
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r2_mgcheck_n$N.log 2>&1
echo "mgcheck rc=$?"; tail -3 gpurun_out/r2_mgcheck_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err
echo "bench rc=$?"; head -c 700 gpurun_out/r2_bench_c3_n$N.json; tail -3 gpurun_out/r2_bench_c3_n$N.err
