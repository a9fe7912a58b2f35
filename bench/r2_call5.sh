#!/bin/bash
# 1-GPU call: LayerNorm-backward revert check, all-gather GEMM on a fake world (local peers), step profile with PDL off.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -n 4 gpurun_out/r2_tests.log
timeout 300 python bench/ln_bench.py > gpurun_out/ln_bench_r2.jsonl 2>&1; cat gpurun_out/ln_bench_r2.jsonl
timeout 300 python bench/ag_gemm_local.py > gpurun_out/ag_gemm_local.jsonl 2> gpurun_out/ag_gemm_local.err; cat gpurun_out/ag_gemm_local.jsonl; tail -n 3 gpurun_out/ag_gemm_local.err
OUT=gpurun_out/bench_r2_ab.jsonl; : > $OUT
ab() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_$tag.err | tail -1 | sed "s/^{/{\"ab\": \"$tag\", /" >> $OUT; tail -1 $OUT | cut -c1-330; }
ab base ZTB_X=0
ab nofold ZTB_FOLD_DROPOUT_BWD=0
ab nopdl ZTB_PDL=0
timeout 600 python bench/profile_step.py > gpurun_out/step_profile.txt 2>&1; head -n 30 gpurun_out/step_profile.txt | cut -c1-160
timeout 600 python bench/gap_analysis.py > gpurun_out/r2_gap.txt 2>&1; head -n 12 gpurun_out/r2_gap.txt | cut -c1-200
