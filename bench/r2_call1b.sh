#!/bin/bash
# Round-2 call 1b (1 GPU): sanitizer passes, isolated-kernel records, ncu captures, 1-GPU sweep of both arms.
mkdir -p gpurun_out
set -x
timeout 600 python bench/attn_bench.py > gpurun_out/attn_bench_r2.jsonl 2> gpurun_out/attn_bench_r2.err; cat gpurun_out/attn_bench_r2.jsonl
bash bench/sweep.sh 1 full; cp gpurun_out/sweep_n1.jsonl gpurun_out/sweep_n1_done.jsonl
bash bench/sanitize.sh
bash bench/capture_profiles.sh > gpurun_out/capture_profiles.log 2>&1; tail -n 5 gpurun_out/capture_profiles.log
