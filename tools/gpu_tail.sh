#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log | cut -c1-250; }
run tests_layers python -m pytest tests/test_gpu_layers.py -q -x -k "tail or head or pool"
TAILN=30 run breakdown python tools/step_breakdown.py 16
