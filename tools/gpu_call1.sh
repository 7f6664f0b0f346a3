#!/bin/bash
# first GPU call: fp32 parity tests + sampler tests + UMMA descriptor probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for am in 0 1; do for bm in 0 1; do for sw in 0 1; do
  timeout 60 ./tools/umma_probe $am $bm $sw >> gpurun_out/probe.txt 2>&1 || echo "probe $am $bm $sw rc=$?" >> gpurun_out/probe.txt
done; done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/probe.txt | grep -E "SUMMARY|rc=|ERROR"
cat gpurun_out/pytest_gpu.txt
