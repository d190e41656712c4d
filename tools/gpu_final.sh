#!/bin/bash
# round artifacts in one call: tests, bench (+reference arm), ncu launch list + full capture of the dominant kernel,
# per-layer conv bench, eigensolver bench, serial step breakdown, parity report, other configs
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-220; }
run tests python -m pytest tests -q -m gpu -x
run bench python bench.py --steps 5 --warmup 3
run bench_ref python bench.py --impl reference --steps 1 --warmup 0
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2
run ncu_full ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -s 20 -c 3 -f -o gpurun_out/prof_conv python tools/profile_step.py 2
TAILN=14 run conv_bench python tools/conv_bench.py 16
TAILN=8 run jacbench python tools/jacobi_bench.py
TAILN=32 run breakdown python tools/step_breakdown.py 16
run parity python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -q -m gpu -s
TAILN=12 run configs python tools/config_runs.py
cat gpurun_out/summary.txt; tail -1 gpurun_out/bench.log | cut -c1-600
