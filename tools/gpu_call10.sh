#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/t10.txt
for d in 0 8000; do echo "delay $d" >> gpurun_out/t10.txt; VMB_PHASE_DELAY=$d timeout 120 python tools/quick_time.py umma 2>&1 | grep fwdbwd >> gpurun_out/t10.txt; done
VMB_PHASE_DELAY=0 VMB_LIB=$PWD/vmap_b200/libvmap_b200_trace.so timeout 300 python tools/trace_umma.py > gpurun_out/trace.txt 2>&1
cat gpurun_out/t10.txt; grep -E "tile 1|coarse|issuer 0" gpurun_out/trace.txt | cut -c1-400
