#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_tests.log
tail -3 gpurun_out/r2t_tests.log
timeout 300 python tools/b1_latency_probe.py > gpurun_out/r2t_latency.txt 2>&1
grep -E "median|device-resident" gpurun_out/r2t_latency.txt
