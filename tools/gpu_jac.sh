#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-250; }
TAILN=8 run jacbench python tools/jacobi_bench.py
run tests python -m pytest tests -q -m gpu -x ${PYTEST_K:+-k "$PYTEST_K"}
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt; tail -1 gpurun_out/bench.log | cut -c1-400
