#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-8} gpurun_out/$name.log | cut -c1-250; }
WCTB_TEST_CONV_IMPLS=6 run tests_pairs python -m pytest tests/test_gpu_layers.py -q -x -k "tensor_core"
if grep -q "tests_pairs rc=0" gpurun_out/summary.txt; then TAILN=16 run conv_bench python tools/conv_bench.py 16; fi
