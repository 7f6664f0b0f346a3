#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/phase.txt
for d in 0 300 600 1000 1500 3000 8000 15000; do
  echo "delay $d" >> gpurun_out/phase.txt
  VMB_PHASE_DELAY=$d timeout 120 python tools/quick_time.py umma 2>&1 | grep fwdbwd >> gpurun_out/phase.txt
done
timeout 600 python -m pytest tests/test_umma_gpu.py -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/phase.txt
cat gpurun_out/phase.txt
