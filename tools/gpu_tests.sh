#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/tests.log | cut -c1-220
