#!/bin/bash
# Round-2 call 1 (1 GPU): kernel numerics, GEMM variants, A/B of the background optimizer update, step profile.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -n 5 gpurun_out/r2_tests.log
timeout 600 python bench/gemm_bench.py > gpurun_out/r2_gemm_bench.jsonl 2> gpurun_out/r2_gemm_bench.err
tail -n 9 gpurun_out/r2_gemm_bench.jsonl
ZTB_UPDATE_MODE=wide timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_wide.err | tail -1 > gpurun_out/r2_bench_wide.json
cat gpurun_out/r2_bench_wide.json
timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_bg.err | tail -1 > gpurun_out/r2_bench_bg.json
cat gpurun_out/r2_bench_bg.json
ZTB_ATTN_DQ_BF16=1 timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_dq16.err | tail -1 > gpurun_out/r2_bench_dq16.json
cat gpurun_out/r2_bench_dq16.json
ZTB_PDL=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "not resume and not train_cli" > gpurun_out/r2_tests_pdl.log 2>&1; echo "pdl tests rc=$?" >> gpurun_out/r2_tests_pdl.log; tail -n 3 gpurun_out/r2_tests_pdl.log
ZTB_PDL=1 timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_pdl.err | tail -1 > gpurun_out/r2_bench_pdl.json
cat gpurun_out/r2_bench_pdl.json
ZTB_LMHEAD_2CTA=1 timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_lm2.err | tail -1 > gpurun_out/r2_bench_lm2.json
cat gpurun_out/r2_bench_lm2.json
ZTB_GEMM_DIRECT=1 timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_direct1.err | tail -1 > gpurun_out/r2_bench_direct1.json
cat gpurun_out/r2_bench_direct1.json
ZTB_CUDA_GRAPH=1 timeout 600 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_bench_graph.err | tail -1 > gpurun_out/r2_bench_graph.json
cat gpurun_out/r2_bench_graph.json; tail -n 3 gpurun_out/r2_bench_graph.err
timeout 600 python bench/gap_analysis.py > gpurun_out/r2_gap.txt 2>&1
tail -n 34 gpurun_out/r2_gap.txt
