#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/tests_pipe.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/tests_pipe.log
for cfg in "--groups 1" "--groups 2" "--groups 4" "--groups 8" "--groups 4 --batch 16" "--groups 2 --batch 4" "--groups 4 --oversub 8"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', '-> fps %.1f  ms/step %.2f  e2e %.1f  conv TF/s %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']))"
done 2>&1 | tee gpurun_out/sweep.log
