#!/bin/bash
mkdir -p gpurun_out
VMB_LIB=$PWD/vmap_b200/libvmap_b200_trace.so timeout 300 python tools/trace_umma.py > gpurun_out/trace.txt 2>&1
cat gpurun_out/trace.txt
