#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_umma_gpu.py -m gpu -q -x -k "golden or oracle_parity or eval_points_tensor" 2>&1 | tail -4 > gpurun_out/k1.txt
for rep in 1 2; do
for lib in libvmap_b200_prev.so libvmap_b200_noearly.so libvmap_b200.so; do
  echo "== $lib" >> gpurun_out/k1.txt
  VMB_LIB=$PWD/vmap_b200/$lib timeout 200 python tools/quick_time.py umma 2>&1 | grep -E "fwdbwd|step" >> gpurun_out/k1.txt
done
done
cat gpurun_out/k1.txt
