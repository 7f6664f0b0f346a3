#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layerwise_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/lw.txt
timeout 200 python tools/lw_time.py >> gpurun_out/lw.txt 2>&1
H=128 R=1200 S=14 timeout 200 python tools/lw_time.py >> gpurun_out/lw.txt 2>&1
GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/lw_launches.csv python tools/lw_time.py > gpurun_out/lw_ncu.log 2>&1
cat gpurun_out/lw.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lw_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:70], r[-1])
PY
