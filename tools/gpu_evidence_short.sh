#!/bin/bash
# Short re-validation after a change that does not touch the fused step kernel:  bash tools/gpu_evidence_short.sh
#   GPU test suite, smoke(), our bench arm (no CPU leg), layer-wise step times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/ev_tests.txt; cat gpurun_out/ev_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.txt 2>&1; tail -3 gpurun_out/ev_smoke.txt
timeout 500 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/ev_bench_ours_nocpu.json 2> gpurun_out/ev_bench_ours.err; tail -2 gpurun_out/ev_bench_ours.err
for shape in "256 4800 32" "256 600 32" "128 1200 14"; do timeout 120 python tools/lw_profile.py $shape; done > gpurun_out/ev_lw_times.txt 2>&1; cat gpurun_out/ev_lw_times.txt
