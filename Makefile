# Convenience targets (the driver uses __graft_entry__.py / pytest / bench.py directly)
PY ?= python

build:            ## libsdcgym.so (nvcc, sm_100a) + CPU oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## everything that runs without a GPU (a few minutes)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity tests proper (B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:
	$(PY) __graft_entry__.py smoke

bench:
	$(PY) bench.py

golden:           ## regenerate tests/golden from the unmodified reference (build container only)
	OPENBLAS_NUM_THREADS=1 $(PY) tests/golden/make_golden.py

.PHONY: build test-cpu test-gpu smoke bench golden
