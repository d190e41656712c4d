#!/bin/bash
# tests + bench + ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run tests python -m pytest tests -q -m gpu -x
run bench python bench.py --steps 5 --warmup 3
run bench_ref python bench.py --impl reference --steps 1 --warmup 0
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2
run ncu_full ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 20 -c 3 -f -o gpurun_out/prof_conv python tools/profile_step.py 2
cat gpurun_out/summary.txt; tail -1 gpurun_out/bench.log
