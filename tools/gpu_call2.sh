#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/pytest_gpu.txt
timeout 300 python tools/quick_time.py fp32 umma > gpurun_out/quick_time.txt 2>&1
tail -60 gpurun_out/pytest_gpu.txt; cat gpurun_out/quick_time.txt
