#!/bin/bash
mkdir -p gpurun_out
for cfg in "--batch 16 --groups 4" "--batch 15 --groups 3" "--batch 30 --groups 2" "--batch 30 --groups 6" "--batch 30 --groups 3" "--batch 32 --groups 4" "--batch 45 --groups 3" "--batch 24 --groups 3"; do
  timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', '-> fps %.1f  ms/step %.2f  e2e %.1f  conv TF/s %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']))"
done 2>&1 | tee gpurun_out/sweep.log
