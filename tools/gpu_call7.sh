#!/bin/bash
mkdir -p gpurun_out
(timeout 120 ./tools/umma_bench 148) > gpurun_out/umma_bench.txt 2>&1
cat gpurun_out/umma_bench.txt
