#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv3_probe.py > gpurun_out/conv3_probe.log 2>&1; echo rc=$?
cat gpurun_out/conv3_probe.log | tail -40
