#!/bin/bash
# Round-2 call 2 (N GPUs, default 2): multi-GPU correctness, comm roofline, bench on both data planes, kernel table.
N=${1:-2}
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests/test_multigpu.py -x -q -m gpu > gpurun_out/r2_multigpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_n$N.log
tail -n 6 gpurun_out/r2_multigpu_n$N.log
timeout 600 $TR --master-port 29611 bench/comm_bench.py > gpurun_out/r2_comm_n$N.log 2>&1; tail -n 2 gpurun_out/r2_comm_n$N.log
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/r2_bench_n$N.err | tail -1 > gpurun_out/r2_bench_n$N.json; cat gpurun_out/r2_bench_n$N.json
ZTB_NVLS=0 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/r2_bench_p2p_n$N.err | tail -1 > gpurun_out/r2_bench_p2p_n$N.json; cat gpurun_out/r2_bench_p2p_n$N.json
timeout 600 $TR --master-port 29614 bench/kernel_table.py > gpurun_out/r2_ktable_n$N.log 2>&1; tail -n 14 gpurun_out/r2_ktable_n$N.log
timeout 300 $TR --master-port 29615 -m zero_transformer_b200.utils.selftest > gpurun_out/selftest_n$N.log 2>&1; tail -n 3 gpurun_out/selftest_n$N.log
