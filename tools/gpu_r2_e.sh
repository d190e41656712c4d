#!/bin/bash
# round-2 GPU call E: full GPU suite after the epilogue rewrite, conv bench, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=10 -p no:cacheprovider > gpurun_out/r2e_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.txt; tail -6 gpurun_out/r2e_pytest.txt
timeout 300 python tools/conv_bench.py 16 > gpurun_out/r2e_convbench.txt 2>&1; cat gpurun_out/r2e_convbench.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'])
print(json.dumps(d['kernel_ms_per_step']))
PY
