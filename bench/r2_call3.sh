#!/bin/bash
# Round-2 call 3 (8 GPUs, expensive: keep it short): comm roofline at n=8 on the default (NVLS) plane, headline bench for both
# arms + small-batch point, kernel table (zero NCCL kernels between step boundaries), selftest, the n=8 correctness tests.
N=${1:-8}
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29711 bench/comm_bench.py > gpurun_out/r2_comm_n$N.log 2>&1; tail -n 2 gpurun_out/r2_comm_n$N.log
bash bench/sweep.sh $N quick
timeout 400 $TR --master-port 29714 bench/kernel_table.py > gpurun_out/r2_ktable_n$N.log 2>&1; tail -n 6 gpurun_out/r2_ktable_n$N.log
timeout 200 $TR --master-port 29715 -m zero_transformer_b200.utils.selftest > gpurun_out/selftest_n$N.log 2>&1; tail -n 2 gpurun_out/selftest_n$N.log
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "(oracle or allgather_gemm or nccl_backend) and 8" > gpurun_out/r2_multigpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_n$N.log
tail -n 4 gpurun_out/r2_multigpu_n$N.log
