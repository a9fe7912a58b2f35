#!/bin/bash
# compute-sanitizer pass over the kernel numerics tests (run under gpurun, 1 GPU).  The reference has no
# sanitizer / race-detection story at all.  memcheck: every elementwise / optimizer / embedding kernel and a
# small instance of each tcgen05 kernel family; racecheck: the shared-memory users among the non-TMA kernels (LayerNorm
# backward incl. its last-block finalize, cross entropy, the Adam kernel's flag prologue).  Cross-DEVICE races are invisible
# to racecheck; those are covered by tests/test_multigpu.py (oracle + random-skew stress) and the memcheck pass of
# bench/sanitize_multigpu.sh over the 2-rank fused update.  The cross-rank kernels themselves (comm.cu pull / reduce-scatter /
# flags, the multi-peer AdamW kernel, the all-gather GEMM) run under memcheck here through the single-GPU fake-world tests.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
K1='layernorm or gelu_kernels or embedding or fused_adamw or dropout_add_fwd or dropout_backward_is_served or comm_kernels_fake_world or reduce_scatter_adamw_fake_world'
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$K1" > gpurun_out/sanitize_memcheck_elementwise.log 2>&1; echo "memcheck elementwise+optim rc=$?" | tee gpurun_out/sanitize_summary.txt
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemm_nt and 256-256-64 or test_gemm_gelu or direct_store and 3 or ksplit and float32 and False or test_attention_fwd and 2-128-2-128 or test_attention_bwd and 1-128-1-128 or bf16_dq and 1-128-1 or allgather_gemm_fake_world" > gpurun_out/sanitize_memcheck_tc.log 2>&1; echo "memcheck tcgen05 rc=$?" | tee -a gpurun_out/sanitize_summary.txt
timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "layernorm and 2048 or cross_entropy and 256 or fused_adamw_background" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_summary.txt
tail -n 4 gpurun_out/sanitize_memcheck_elementwise.log gpurun_out/sanitize_memcheck_tc.log gpurun_out/sanitize_racecheck.log
