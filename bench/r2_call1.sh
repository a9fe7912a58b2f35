#!/bin/bash
# Round-2 comprehensive 1-GPU call: full GPU test suite, micro-benchmarks, A/B matrix of the step switches, the
# BASELINE-config sweep of both arms, sanitizer passes, ncu captures, step profile / gap analysis.
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -n 25 gpurun_out/r2_tests.log
timeout 300 python bench/ln_bench.py > gpurun_out/ln_bench_r2.jsonl 2>&1; cat gpurun_out/ln_bench_r2.jsonl
OUT=gpurun_out/bench_r2_ab.jsonl; : > $OUT
ab() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_$tag.err | tail -1 | sed "s/^{/{\"ab\": \"$tag\", /" >> $OUT; tail -1 $OUT | cut -c1-330; }
ab base ZTB_X=0
ab nofold ZTB_FOLD_DROPOUT_BWD=0
ab dq32 ZTB_ATTN_DQ_BF16=0
ab nopdl ZTB_PDL=0
ab dq32_nopdl ZTB_ATTN_DQ_BF16=0 ZTB_PDL=0
ab carveout_all ZTB_CARVEOUT_ALL=1
ab wide ZTB_UPDATE_MODE=wide
ab graph ZTB_CUDA_GRAPH=1
ab base2 ZTB_X=0
timeout 600 python bench/gap_analysis.py > gpurun_out/r2_gap.txt 2>&1; head -n 12 gpurun_out/r2_gap.txt | cut -c1-200
timeout 600 python bench/profile_step.py > gpurun_out/step_profile.txt 2>&1; head -n 24 gpurun_out/step_profile.txt | cut -c1-160
timeout 600 python bench/attn_bench.py > gpurun_out/attn_bench_r2.jsonl 2> gpurun_out/attn_bench_r2.err; cat gpurun_out/attn_bench_r2.jsonl
timeout 600 python bench/gemm_bench.py > gpurun_out/gemm_bench_r2.jsonl 2> gpurun_out/gemm_bench_r2.err
bash bench/sweep.sh 1 full
bash bench/sanitize.sh
bash bench/capture_profiles.sh > gpurun_out/capture_profiles.log 2>&1; tail -n 5 gpurun_out/capture_profiles.log
