#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt; tail -4 gpurun_out/r2m_pytest.txt
timeout 300 python tools/cov_bench.py 16 2>&1 | tee gpurun_out/r2m_cov_bench.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'])"
