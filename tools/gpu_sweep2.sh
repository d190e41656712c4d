#!/bin/bash
mkdir -p gpurun_out
for cfg in "--oversub 4" "--oversub 2" "--oversub 8" "--oversub 1" "--no-prio" "--batch 45 --groups 3" "--batch 60 --groups 4"; do
  timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', '-> fps %.1f  ms/step %.2f  e2e %.1f  conv TF/s %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']))"
done 2>&1 | tee gpurun_out/sweep2.log
