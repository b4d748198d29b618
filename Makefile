# convenience targets (the driver uses __graft_entry__.py, pytest and bench.py directly)
PY ?= python
.PHONY: build test test-gpu bench bench-ref golden clean
build:            ## libiexa_b200.so (nvcc, sm_100a) + oracle + host-check library + NVRTC cross-compile check
	$(PY) __graft_entry__.py
test: build       ## CPU suite
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu: build   ## needs a B200
	$(PY) -m pytest tests -q -m gpu
bench: build      ## one JSON line (needs a B200)
	$(PY) bench.py
bench-ref:        ## CPU restatement of the reference's evaluator, all host threads
	$(PY) bench.py --impl reference
golden:           ## regenerate the independent sympy known-answer fixtures
	$(PY) tests/golden/make_sympy_golden.py
clean:
	rm -rf infiniteexamodels.jl_b200/build infiniteexamodels.jl_b200/libiexa_b200.so oracle/liboracle.so tests/hostcheck/libiexa_hostcheck.so
