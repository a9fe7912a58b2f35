#!/bin/bash
# compute-sanitizer pass over the kernel numerics tests (run under gpurun, 1 GPU).  The reference has no
# sanitizer / race-detection story at all; cross-device races cannot be seen by racecheck,
# those are covered by the 2-GPU protocol tests instead.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
K1='layernorm or gelu_kernels or embedding or fused_adamw or dropout_add_fwd'
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$K1" > gpurun_out/sanitize_memcheck_elementwise.log 2>&1; echo "memcheck elementwise rc=$?" | tee gpurun_out/sanitize_summary.txt
timeout 500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemm_nt and 256-256-64 or test_gemm_gelu or test_attention_fwd and 2-128-2-128 or test_attention_bwd and 1-128-1-128" > gpurun_out/sanitize_memcheck_tc.log 2>&1; echo "memcheck tcgen05 rc=$?" | tee -a gpurun_out/sanitize_summary.txt
timeout 400 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "layernorm and 2048 or cross_entropy and 256" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_summary.txt
tail -n 4 gpurun_out/sanitize_memcheck_elementwise.log gpurun_out/sanitize_memcheck_tc.log gpurun_out/sanitize_racecheck.log
