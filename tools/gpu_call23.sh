#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_sampler_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/ingest.txt
cat gpurun_out/ingest.txt
