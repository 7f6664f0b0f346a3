#!/bin/bash
# Round-end style validation on one B200 (run through gpurun):  bash tools/gpu_validate.sh [quick]
#   full GPU test suite, smoke(), both bench arms; writes everything under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/validate_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/validate_smoke.txt 2>&1
if [ "$1" != "quick" ]; then
  timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/validate_tests.txt; tail -2 gpurun_out/validate_smoke.txt; cat gpurun_out/bench_ref.json 2>/dev/null | cut -c1-600; cat gpurun_out/bench_ours.json | cut -c1-1500
