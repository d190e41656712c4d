#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -2
for cfg in "--batch 8 --groups 4" "--batch 8 --groups 4" "--batch 16 --groups 4" "--batch 30 --groups 3" "--batch 30 --groups 2" "--batch 16 --groups 4"; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', '-> fps %.1f  ms/step %.2f  e2e %.1f (%.2f ms)'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done 2>&1 | tee gpurun_out/stab.log
