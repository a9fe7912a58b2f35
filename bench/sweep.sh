#!/bin/bash
# Throughput sweep for the BASELINE.md table: run under `gpurun --gpus N -- bash bench/sweep.sh N`.
# Writes one JSON line per (model, arm) into gpurun_out/sweep_n${N}.jsonl.  Every line is a normal bench.py record.
set -u
N=${1:-1}
OUT=gpurun_out/sweep_n${N}.jsonl
mkdir -p gpurun_out; : > "$OUT"
run() {  # env assignments..., then bench args
  if [ "$N" -gt 1 ]; then
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus "$N" --steps 6 --warmup 3 ${ARM:+--impl $ARM} 2>/dev/null | tail -1 >> "$OUT"
  else
    env "$@" timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 ${ARM:+--impl $ARM} 2>/dev/null | tail -1 >> "$OUT"
  fi
}
ARM=
run ZTB_BENCH_MODEL=417m
run ZTB_BENCH_MODEL=760m
run ZTB_BENCH_MODEL=1_3b
run ZTB_BENCH_MODEL=1_3b ZTB_BENCH_SEQ=2048 ZTB_BENCH_ROWS=64 ZTB_BENCH_ACCUM=8     # BASELINE config #5 shape: 8 micro-batches of 8 x 2048 tokens per GPU per step
run ZTB_BENCH_MODEL=1_3b ZTB_NVLS=0                                                 # P2P data plane
ARM=baseline
run ZTB_BENCH_MODEL=1_3b                                                            # stock PyTorch + NCCL arm
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except ValueError:
        continue
    c = d.get("config", {})
    print(f"{d.get('impl','?'):9s} {str(c.get('model'))[:28]:28s} seq {c.get('seq_len')} accum {c.get('grad_accum')}  n={d.get('n_gpus')}  "
          f"{d.get('value', 0) / 1e3:8.1f}k tok/s  {d.get('ms_per_step')} ms  e2e {d.get('e2e', {}).get('value', 0) / 1e3:.1f}k")
PY
