#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/full_gpu.txt
GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/lw_launches.csv python tools/lw_time.py > gpurun_out/lw_ncu.log 2>&1
cat gpurun_out/full_gpu.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lw_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:60], r[-1])
PY
