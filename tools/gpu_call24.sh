#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/ingest.txt
timeout 300 python tools/ingest_time.py > gpurun_out/ingest_time.json 2> gpurun_out/ingest_time.err
cat gpurun_out/ingest.txt; cat gpurun_out/ingest_time.json; tail -5 gpurun_out/ingest_time.err
