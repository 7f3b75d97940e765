#!/bin/bash
# lv.sh TAG [LIB]: per-kernel durations of one 32768-bead CCpol gradient pass -> gpurun_out/lv_TAG.csv (printed)
TAG=$1; [ -n "$2" ] && export PIMDK_LIB=$2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccpol_ -s 8 -c 17 --csv --log-file gpurun_out/lv_$TAG.csv python tools/prof_ccpol.py ${MODE:-0} 32768 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/lv_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
st=[i for i,r in enumerate(rows) if 'setup' in r[4]]
rows=rows[st[0]:(st[1] if len(st)>1 else len(rows))]   # one whole pass (7 or 8 kernels)
tot=0
for r in rows:
    print('$TAG', r[4].split('::')[-1][:28], r[8], int(r[-1])/1e6); tot+=int(r[-1])
print('$TAG total ms', tot/1e6)
PY
