#!/bin/bash
# Round-2 multi-GPU call (8 GPUs of one box; expensive, keep it tight): correctness of both data planes at world 2 and 8,
# communication-kernel roofline at n = 2 and 8, scaling points of the headline config for both arms, kernel table
# (zero NCCL kernels between step boundaries), selftest.
mkdir -p gpurun_out
set -x
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
timeout 1500 python -m pytest 'tests/test_multigpu.py::test_fused_backend_matches_nccl_backend[1-8]' 'tests/test_multigpu.py::test_fused_backend_matches_nccl_backend[0-8]' 'tests/test_multigpu.py::test_fused_backend_matches_nccl_backend[1-2]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[1-2]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[1-8]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[0-2]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[0-8]' 'tests/test_multigpu.py::test_allgather_gemm_matches_plain_gemm[2]' 'tests/test_multigpu.py::test_allgather_gemm_matches_plain_gemm[8]' 'tests/test_multigpu.py::test_flag_protocol_survives_random_skew[1-2]' 'tests/test_multigpu.py::test_flag_protocol_survives_random_skew[1-8]' 'tests/test_multigpu.py::test_flag_protocol_survives_random_skew[0-2]' 'tests/test_multigpu.py::test_selftest_probe_two_gpus' -q -m gpu > gpurun_out/r2_multigpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu.log
tail -n 8 gpurun_out/r2_multigpu.log
for n in 8 2; do
  timeout 400 bash -c "$(declare -f TR); TR $n --master-port 2971$n bench/comm_bench.py" > gpurun_out/r2_comm_n$n.log 2>&1; tail -n 1 gpurun_out/r2_comm_n$n.log | cut -c1-600
done
bash bench/sweep.sh 8 quick
OUT=gpurun_out/sweep_scaling.jsonl; : > $OUT
for n in 2 4; do
  timeout 600 bash -c "$(declare -f TR); TR $n --master-port 2973$n bench.py --gpus $n --steps 6 --warmup 3" 2>>gpurun_out/sweep_scaling.err | tail -1 >> $OUT
  timeout 600 bash -c "$(declare -f TR); TR $n --master-port 2974$n bench.py --gpus $n --steps 6 --warmup 3 --impl baseline" 2>>gpurun_out/sweep_scaling.err | tail -1 >> $OUT
done
cut -c1-260 $OUT
timeout 400 bash -c "$(declare -f TR); TR 8 --master-port 29751 bench/kernel_table.py" > gpurun_out/r2_ktable_n8.log 2>&1; tail -n 6 gpurun_out/r2_ktable_n8.log | cut -c1-300
timeout 200 bash -c "$(declare -f TR); TR 8 --master-port 29752 -m zero_transformer_b200.utils.selftest" > gpurun_out/selftest_n8.log 2>&1; tail -n 2 gpurun_out/selftest_n8.log | cut -c1-300
