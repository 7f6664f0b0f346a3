#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 30 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
VMB_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_umma -s 4 -c 1 -o gpurun_out/prof_umma_r01 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_ours.json | cut -c1-2500; cat gpurun_out/bench_ref.json | cut -c1-600; tail -3 gpurun_out/smoke.txt; grep -E "k_step|k_adam|k_mask" gpurun_out/launches_r01.csv | head -6 | cut -c60-400
