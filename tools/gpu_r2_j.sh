#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2j_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.txt; tail -4 gpurun_out/r2j_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'])
print(json.dumps(d['covariance_hbm']))
print(json.dumps(d['hbm_gbs_by_stage']))
PY
