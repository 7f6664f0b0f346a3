#!/bin/bash
# compute-sanitizer passes over the hot-path kernels (run through gpurun, ONE GPU):  bash tools/gpu_sanitize.sh
#   memcheck + racecheck of small K1 (tcgen05 and fp32 flavours) / K2 / K3 / K4 invocations taken from the GPU test suite.
mkdir -p gpurun_out
SEL_K1="tests/test_step_gpu.py tests/test_umma_gpu.py::test_golden_render_umma tests/test_umma_gpu.py::test_umma_training_tracks_fp32_kernel"
SEL_K34="tests/test_sampler_gpu.py::test_injected_randoms_reproduce_the_reference tests/test_ingest_gpu.py::test_ingest_many_ids_unknown_and_out_of_range"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL_K1 $SEL_K34 -q -x -p no:cacheprovider \
      > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?" >> gpurun_out/sanitize_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_$tool.txt | tail -8
done
