#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cov_probe.py > gpurun_out/cov_probe.log 2>&1; echo rc=$?; cat gpurun_out/cov_probe.log | tail -40
