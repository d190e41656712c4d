#!/bin/bash
# quick loop: transform+layer tests, bench (no cpu baseline), launch list
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
run tests python -m pytest tests -q -m gpu -x ${PYTEST_K:+-k "$PYTEST_K"}
run bench python bench.py --steps 5 --warmup 3 --batch ${BATCH:-8} --no-cpu-baseline
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2
cat gpurun_out/summary.txt; tail -1 gpurun_out/bench.log | cut -c1-300
