#!/bin/bash
# round-2 GPU call H: eigensolver after the half-warp split; predicted-convergence level; batch / group sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -k "jacobi or rank or wct_level or golden" -p no:cacheprovider > gpurun_out/r2h_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.txt; tail -4 gpurun_out/r2h_pytest.txt
timeout 300 python tools/jacobi_bench.py > gpurun_out/r2h_jacobi_bench.txt 2>&1; cat gpurun_out/r2h_jacobi_bench.txt
run() { timeout 300 python bench.py --steps 4 --warmup 3 --no-roofline --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-44s value %.1f  e2e %.1f  ms/step %.1f' % ('$*', d['value'], d['e2e']['value'], d['ms_per_step']))"; }
{
run --batch 30 --groups 2
run --batch 30 --groups 2 --jacobi-tolq 3e-4
run --batch 30 --groups 2 --jacobi-tolq 1e-3
run --batch 30 --groups 3
run --batch 45 --groups 3
run --batch 60 --groups 4
run --batch 30 --groups 2 --no-prio
run --batch 30 --groups 2 --oversub 2
} | tee gpurun_out/r2h_sweep.txt
