#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/lw_time.py > gpurun_out/lw_time.txt 2>&1
H=128 R=1200 S=14 timeout 200 python tools/lw_time.py >> gpurun_out/lw_time.txt 2>&1
GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 45 --csv --log-file gpurun_out/lw_launches.csv python tools/lw_time.py > gpurun_out/ncu_lw.log 2>&1
cat gpurun_out/lw_time.txt; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lw_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows[:40]:
    name=r[4].split('(')[0][:60]; ns=float(r[-1]); tot+=ns
    print(f"{name:62s} {r[7]:>14s} {ns/1e3:9.1f} us")
print("sum", tot/1e3, "us over", min(40,len(rows)), "launches")
PY
