#!/bin/bash
# Round-2 call for a 2- or 4-GPU box: `gpurun --gpus N -- bash bench/r2_call_n.sh N`.  Multi-GPU tests at world N, comm
# roofline at n = N, the BASELINE-config sweep of both arms at n = N, selftest.
N=${1:-2}
MODE=${2:-full}
mkdir -p gpurun_out
set -x
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -k "[$N] or -$N] or selftest" > gpurun_out/r2_multigpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_n$N.log
tail -n 8 gpurun_out/r2_multigpu_n$N.log
timeout 300 bash -c "$(declare -f TR); TR $N --master-port 2971$N bench/comm_bench.py" > gpurun_out/r2_comm_n$N.log 2>&1; tail -n 1 gpurun_out/r2_comm_n$N.log | cut -c1-900
bash bench/sweep.sh $N $MODE
timeout 200 bash -c "$(declare -f TR); TR $N --master-port 29752 -m zero_transformer_b200.utils.selftest" > gpurun_out/selftest_n$N.log 2>&1; tail -n 2 gpurun_out/selftest_n$N.log | cut -c1-300
