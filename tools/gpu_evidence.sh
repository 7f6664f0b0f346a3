#!/bin/bash
# Round-end evidence on one B200 (run through gpurun):  bash tools/gpu_evidence.sh
#   GPU test suite, smoke(), both bench arms, launch list, one full ncu capture of the step kernel, cycle trace,
#   layer-wise path timings + launch lists.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/ev_tests.txt; cat gpurun_out/ev_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.txt 2>&1; tail -3 gpurun_out/ev_smoke.txt
timeout 500 python bench.py --steps 200 --warmup 10 > gpurun_out/ev_bench_ours.json 2> gpurun_out/ev_bench_ours.err; tail -2 gpurun_out/ev_bench_ours.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/ev_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-extras > gpurun_out/ev_ncu_bench.log 2>&1
VMB_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 4 -c 1 \
    -o gpurun_out/prof_k_step_fused -f python bench.py --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/ev_ncu_full.log 2>&1
tail -3 gpurun_out/ev_ncu_full.log
VMB_LIB=vmap_b200/libvmap_b200_trace.so timeout 200 python tools/trace_fused.py > gpurun_out/ev_trace.txt 2>&1; tail -3 gpurun_out/ev_trace.txt
for shape in "256 4800 32" "256 600 32" "128 1200 14"; do timeout 120 python tools/lw_profile.py $shape; done > gpurun_out/ev_lw_times.txt 2>&1; cat gpurun_out/ev_lw_times.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_lw_launches_4800.csv \
    python tools/lw_profile.py 256 4800 32 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_lw_launches_bg.csv \
    python tools/lw_profile.py 128 1200 14 > /dev/null 2>&1
ls -la gpurun_out | grep -E "ev_|prof_k_step_fused"
