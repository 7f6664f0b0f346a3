#!/bin/bash
# ncu recipes used for profiles/ (run through gpurun, ONE GPU):  bash tools/gpu_profile.sh {k1|launches|layerwise}
mkdir -p gpurun_out
case "$1" in
  k1)        # one full capture of the fused step kernel (forward+backward), BASELINE cfg 2
    VMB_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 4 -c 1 \
      -o gpurun_out/prof_k_step_fused -f python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1 ;;
  launches)  # launch list of a short bench run (per-kernel durations; cold cache, serialised)
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1 ;;
  layerwise) # launch list + full capture of the weight-stationary GEMMs of the wide-model path (H=256, 4800 x 32)
    GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv \
      --log-file gpurun_out/lw_launches.csv python tools/lw_time.py > gpurun_out/lw_ncu.log 2>&1
    GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -s 20 -c 12 \
      -o gpurun_out/ws_full -f python tools/lw_time.py > gpurun_out/ws_ncu.log 2>&1 ;;
  *) echo "usage: $0 {k1|launches|layerwise}"; exit 2 ;;
esac
ls -la gpurun_out | tail -5
