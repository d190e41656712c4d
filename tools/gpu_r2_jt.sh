#!/bin/bash
# tensor-core eigensolver (k_jacobi_tc): correctness, then timing vs k_jacobi<512>
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -k "jacobi or rank" -p no:cacheprovider > gpurun_out/r2jt_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2jt_pytest.txt; tail -12 gpurun_out/r2jt_pytest.txt
export JAC_CASES="512:1024:15,512:1024:4,512:300:4"
for impl in 2 1; do
  echo "== impl $impl"
  JAC_IMPL=$impl timeout 200 python tools/jacobi_bench.py 2>&1 | tail -4
done | tee gpurun_out/r2jt_jacobi_bench.txt
python -c "
import sys; sys.path.insert(0,'.')
from wct_tf_b200 import _capi
lib=_capi.load(); print('device check rc', lib.wctb200_check_device(None), lib.wctb200_last_error())"
