.PHONY: build test test-gpu bench style
build:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests -q -m "not gpu"
test-gpu:
	python -m pytest tests -q -m gpu
bench:
	python bench.py --gpus 1 --steps 8 --warmup 3
style:
	python -m compileall -q zero_transformer_b200 tests bench.py
sanitize:
	bash bench/sanitize.sh
