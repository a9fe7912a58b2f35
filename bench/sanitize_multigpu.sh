#!/bin/bash
# compute-sanitizer memcheck over the cross-rank kernels (optim.cu NVLS / P2P update, comm.cu signal / wait / pull, the
# all-gather GEMM): the 2-rank fp32-oracle test and the all-gather-GEMM test, every rank under the sanitizer
# (run under gpurun --gpus 2).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --target-processes all --error-exitcode 9 python -m pytest tests/test_multigpu.py -x -q -m gpu \
  -k "fp32_oracle and 2 or allgather_gemm and 2" > gpurun_out/sanitize_memcheck_multigpu.log 2>&1
echo "memcheck multigpu rc=$?" | tee gpurun_out/sanitize_multigpu_summary.txt
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_multigpu.log | tail -n 8
