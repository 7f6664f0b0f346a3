#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/quick_time.py umma 2>&1 | grep -E "fwdbwd|step" > gpurun_out/t13.txt
TRACE_PER=56 VMB_LIB=$PWD/vmap_b200/libvmap_b200_trace.so timeout 300 python tools/trace_umma.py > gpurun_out/trace.txt 2>&1
tail -4 gpurun_out/pytest_gpu.txt; cat gpurun_out/t13.txt; grep -E "tile 1|coarse" gpurun_out/trace.txt | cut -c1-400
