#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_umma_gpu.py -m gpu -q -x -k "golden or oracle_parity" 2>&1 | tail -4 > gpurun_out/k1.txt
timeout 200 python tools/quick_time.py umma >> gpurun_out/k1.txt 2>&1
cat gpurun_out/k1.txt
