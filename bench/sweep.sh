#!/bin/bash
# Throughput sweep for the BASELINE.md table: run under `gpurun --gpus N -- bash bench/sweep.sh N [quick]`.
# Writes one JSON line per (config, arm) into gpurun_out/sweep_n${N}.jsonl.  Every line is a normal bench.py record
# (CUDA-event timed, max over ranks, clocks sampled).  "quick" = headline + small-batch point only (8-GPU budget).
set -u
N=${1:-1}
MODE=${2:-full}
OUT=gpurun_out/sweep_n${N}.jsonl
mkdir -p gpurun_out; : > "$OUT"
run() {  # env assignments..., then bench args
  if [ "$N" -gt 1 ]; then
    env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus "$N" --steps 6 --warmup 3 ${ARM:+--impl $ARM} 2>>gpurun_out/sweep_n${N}.err | tail -1 >> "$OUT"
  else
    env "$@" timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 ${ARM:+--impl $ARM} 2>>gpurun_out/sweep_n${N}.err | tail -1 >> "$OUT"
  fi
}
for ARM in "" baseline; do
  run ZTB_BENCH_MODEL=1_3b                                                            # headline (BASELINE config #4)
  run ZTB_BENCH_MODEL=1_3b ZTB_BENCH_ROWS=8 ZTB_BENCH_ACCUM=1                         # small batch: 8k tokens / GPU / step, communication-visible
  if [ "$MODE" = "full" ]; then
    run ZTB_BENCH_MODEL=417m                                                          # config #2
    run ZTB_BENCH_MODEL=760m                                                          # config #3
    run ZTB_BENCH_MODEL=1_3b ZTB_BENCH_SEQ=2048 ZTB_BENCH_ROWS=64 ZTB_BENCH_ACCUM=8   # config #5 shape: 8 micro-batches of 8 x 2048 tokens
  fi
done
ARM=
if [ "$N" -gt 1 ]; then run ZTB_BENCH_MODEL=1_3b ZTB_NVLS=0; fi                        # P2P data plane (all-gather inside the QKV GEMM)
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except ValueError:
        continue
    c = d.get("config", {})
    print(f"{d.get('impl','?'):9s} {str(c.get('model'))[:28]:28s} seq {c.get('seq_len')} rows {c.get('global_batch')} accum {c.get('grad_accum')}  n={d.get('n_gpus')}  "
          f"{d.get('value', 0) / 1e3:8.1f}k tok/s  {d.get('ms_per_step')} ms  e2e {d.get('e2e', {}).get('value', 0) / 1e3:.1f}k  clocks {d.get('clocks', {}).get('sm_mhz')}")
PY
