#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -x 2>&1 | grep -E "assert|Error|error|FAILED|passed|failed" | head -20 > gpurun_out/t14.txt
cat gpurun_out/t14.txt
