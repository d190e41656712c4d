#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py 30 2 1 > gpurun_out/r2tl_30_2.txt 2>&1; grep -v "^  g" gpurun_out/r2tl_30_2.txt | head -40
timeout 600 python tools/sched_sweep.py 2>&1 | tail -12
