#!/bin/bash
mkdir -p gpurun_out
for cfg in "--groups 2 --prio" "--groups 4 --prio" "--groups 2 --prio --oversub 1" "--groups 2 --oversub 1" "--groups 4 --prio --batch 16"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', '-> fps %.1f  ms/step %.2f  e2e %.1f  conv TF/s %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']))"
done 2>&1 | tee gpurun_out/sweep.log
