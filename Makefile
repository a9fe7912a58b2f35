.PHONY: build test test-gpu test-multigpu bench bench-baseline comm-bench sweep tables sass selftest style sanitize
build:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests -q -m "not gpu"
test-gpu:
	python -m pytest tests -q -m gpu
test-multigpu:          # needs >= 2 GPUs; world 4 / 8 cases skip themselves when the box is smaller
	python -m pytest tests/test_multigpu.py -q -m gpu
bench:
	python bench.py --gpus 1 --steps 8 --warmup 3
bench-baseline:         # stock-library arm (cuBLAS + flash_attn + fused AdamW + NCCL), same shape
	python bench.py --impl baseline --gpus 1 --steps 8 --warmup 3
comm-bench:             # N=2,4,8: reduce-scatter / all-gather / fused-kernel roofline fractions
	python -m torch.distributed.run --nproc-per-node $${N:-2} --master-addr 127.0.0.1 bench/comm_bench.py
sweep:
	bash bench/sweep.sh $${N:-1} full
tables:                 # profiles/*.jsonl -> profiles/RESULTS.md + profiles/baseline_arm.json
	python bench/make_tables.py
sass:                   # SASS mnemonic counts + raw excerpts of the hot kernels (no GPU needed)
	python bench/sass_table.py
selftest:
	python -m torch.distributed.run --nproc-per-node $${N:-2} --master-addr 127.0.0.1 -m zero_transformer_b200.utils.selftest
style:
	python -m compileall -q zero_transformer_b200 tests bench.py
sanitize:
	bash bench/sanitize.sh
