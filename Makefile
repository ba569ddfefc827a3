# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu

bench: build
	$(PY) bench.py

sweep: build
	$(PY) scripts/sweep.py --full-job 700000

golden:
	$(PY) tests/golden/make_golden.py
	$(PY) tests/golden/make_reference_golden.py

clean:
	$(MAKE) -C textreact_b200/csrc clean
	$(MAKE) -C oracle clean

.PHONY: build test-cpu test-gpu bench sweep golden clean
