#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/rank_probe.py > gpurun_out/rank_probe.log 2>&1; echo rc=$?; cat gpurun_out/rank_probe.log | tail -12
