# Convenience targets; the real build lives in hpc_multigpu_matrixmult_b200/Makefile and oracle/Makefile.
#   make            libphpc_b200.so, MPI shim, bin/{main.out,mpirun,fp64_peak,umma_rate}, the CPU checker (+ oracle/_ref if /root/reference exists)
#   make test       CPU test suite (no GPU needed)
#   make test-gpu   parity tests on a B200
#   make bench      SUMMA N=32768 on one B200 (torchrun for 2/4/8, see bench.py)
PY ?= python

all:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: all
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: all
	$(PY) -m pytest tests -q -m gpu

bench: all
	$(PY) bench.py

clean:
	$(MAKE) -C hpc_multigpu_matrixmult_b200 clean
	$(MAKE) -C oracle clean

.PHONY: all test test-gpu bench clean
