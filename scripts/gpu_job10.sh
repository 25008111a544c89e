#!/bin/bash
mkdir -p gpurun_out
KBA_LAUNCH_CHECK=1 timeout 300 tests/cpp/test_facade gpu > gpurun_out/r2_facade10.log 2>&1; echo "facade rc=$?"; tail -20 gpurun_out/r2_facade10.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t10.log
tail -30 gpurun_out/r2_t10.log | cut -c1-600
