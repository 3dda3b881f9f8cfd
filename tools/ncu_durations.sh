#!/bin/bash
# kernel durations of the baseline pass measured by ncu (light metrics), printed as a table
python tools/prof_pass.py --passes 3 > gpurun_out/pp.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"user_sum|user_avg|item_tiled|predict_mae" -s 5 -c 5 --csv --log-file gpurun_out/launches_x.csv python tools/prof_pass.py --passes 3 > gpurun_out/ncu_x.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_x.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[h]; ki=H.index('Kernel Name'); mi=H.index('Metric Name'); vi=H.index('Metric Value')
cur={}
for r in rows[h+1:]:
    name=r[ki].split('(')[0].split('::')[-1]
    cur.setdefault((r[0],name),{})[r[mi].split('.')[0][-14:]]=r[vi]
tot=0
for (i,n),m in cur.items():
    print(i,n,m); tot+=float(m['_time_duration'])
print('sum of kernel durations (us):', tot/1000)
PY
