#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layerwise_gpu.py -m gpu -q -x -s -k "eval_points" 2>&1 | grep -E "passed|failed|FAILED|assert|Error|eval_points" | tail -12 > gpurun_out/ev2.txt
cat gpurun_out/ev2.txt
