#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py 30 2 1 > gpurun_out/r2tl_30_2.txt 2>&1; cat gpurun_out/r2tl_30_2.txt
timeout 300 python tools/timeline.py 30 3 1 > gpurun_out/r2tl_30_3.txt 2>&1; head -8 gpurun_out/r2tl_30_3.txt; tail -3 gpurun_out/r2tl_30_3.txt
timeout 300 python tools/timeline.py 30 2 0 > gpurun_out/r2tl_30_2_noprio.txt 2>&1; head -3 gpurun_out/r2tl_30_2_noprio.txt; tail -3 gpurun_out/r2tl_30_2_noprio.txt
