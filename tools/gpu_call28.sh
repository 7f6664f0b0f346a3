#!/bin/bash
mkdir -p gpurun_out
GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -s 20 -c 12 -o gpurun_out/ws_full -f python tools/lw_time.py > gpurun_out/ws_ncu.log 2>&1
tail -3 gpurun_out/ws_ncu.log
ls -la gpurun_out/ws_full.ncu-rep
