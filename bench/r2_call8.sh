#!/bin/bash
# Round-2 8-GPU call (charged 8x: keep it tight): correctness of both data planes at world 8, communication-kernel
# roofline at n = 8, the BASELINE-config sweep of both arms at n = 8, kernel table of one step (zero NCCL kernels between
# step boundaries), selftest.  World 2 / 4 runs live in bench/r2_call_n.sh (cheaper boxes).
mkdir -p gpurun_out
set -x
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
timeout 900 python -m pytest 'tests/test_multigpu.py::test_fused_backend_matches_nccl_backend[1-8]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[1-8]' 'tests/test_multigpu.py::test_fused_update_matches_fp32_oracle[0-8]' 'tests/test_multigpu.py::test_allgather_gemm_matches_plain_gemm[8]' 'tests/test_multigpu.py::test_flag_protocol_survives_random_skew[1-8]' -q -m gpu > gpurun_out/r2_multigpu_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_n8.log
tail -n 8 gpurun_out/r2_multigpu_n8.log
timeout 300 bash -c "$(declare -f TR); TR 8 --master-port 29718 bench/comm_bench.py" > gpurun_out/r2_comm_n8.log 2>&1; tail -n 1 gpurun_out/r2_comm_n8.log | cut -c1-900
bash bench/sweep.sh 8 full
timeout 300 bash -c "$(declare -f TR); TR 8 --master-port 29751 bench/kernel_table.py" > gpurun_out/r2_ktable_n8.log 2>&1; tail -n 6 gpurun_out/r2_ktable_n8.log | cut -c1-300
timeout 200 bash -c "$(declare -f TR); TR 8 --master-port 29752 -m zero_transformer_b200.utils.selftest" > gpurun_out/selftest_n8.log 2>&1; tail -n 2 gpurun_out/selftest_n8.log | cut -c1-300
