#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag160.py > gpurun_out/diag160.txt 2>&1
timeout 600 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -k cfg0 2>&1 | grep -E "assert|Error|passed|failed|^E" | head -12 >> gpurun_out/diag160.txt
cat gpurun_out/diag160.txt
