#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/mgpu_check.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2>&1
tail -5 gpurun_out/mgpu_check.txt; cat gpurun_out/bench_2gpu.json | cut -c1-700; tail -3 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_ref_2gpu.json | tail -1 | cut -c1-200
