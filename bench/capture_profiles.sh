#!/bin/bash
# Run under gpurun (1 GPU): ncu --set full captures of the hot kernels into gpurun_out/.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm2?_tcgen05 -s 10 -c 5 -f -o gpurun_out/prof_gemm python bench/prof_kernels.py gemm > gpurun_out/prof_gemm.log 2>&1
$NCU -k regex:attn_ -s 8 -c 4 -f -o gpurun_out/prof_attn python bench/prof_kernels.py attn > gpurun_out/prof_attn.log 2>&1
$NCU -k regex:"adamw|layernorm|ce_" -s 5 -c 5 -f -o gpurun_out/prof_misc python bench/prof_kernels.py misc > gpurun_out/prof_misc.log 2>&1
python bench/profile_step.py > gpurun_out/step_profile.txt 2>&1
python - <<'PY' > gpurun_out/cublas_kernel_names.txt 2>&1
import torch
from torch.profiler import profile, ProfilerActivity
a=torch.randn(16384,2048,device='cuda').bfloat16(); b=torch.randn(6144,2048,device='cuda').bfloat16()
for _ in range(3): a@b.t()
with profile(activities=[ProfilerActivity.CUDA]) as p:
    a@b.t(); (a@b.t()).t()@a; torch.cuda.synchronize()
for e in p.key_averages(): print(e.key, e.device_time_total)
PY
tail -n 3 gpurun_out/prof_gemm.log
