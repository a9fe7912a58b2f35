#!/bin/bash
# 2-GPU call: all-gather GEMM variants over real NVLink (+ the fake-world harness on GPU 0) and the world-2 AG test.
mkdir -p gpurun_out
: > gpurun_out/ag_variants.jsonl
for v in 1 2; do
  ZTB_AG_VARIANT=$v FAKE_WORLD=2,8 python bench/ag_gemm_local.py 2>/dev/null | sed "s/^{/{\"variant\": $v, /" >> gpurun_out/ag_variants.jsonl
  ZTB_AG_VARIANT=$v COMM_ONLY_AG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2977$v bench/comm_bench.py 2>/dev/null | tail -n 1 >> gpurun_out/ag_variants.jsonl
done
cut -c1-700 gpurun_out/ag_variants.jsonl
timeout 300 python -m pytest tests/test_multigpu.py -q -m gpu -k "allgather_gemm and 2" 2>&1 | tail -n 2
