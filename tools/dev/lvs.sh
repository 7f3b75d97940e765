#!/bin/bash
# lvs.sh KERNEL_REGEX NAME...: duration of matching kernels for each variant library tools/dev/variants/libpimdk_NAME.so
RX=$1; shift
for n in "$@"; do
  PIMDK_LIB=tools/dev/variants/libpimdk_$n.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$RX -s 2 -c 2 --csv --log-file gpurun_out/lvs_$n.csv python tools/dev/prof_ccpol.py 0 32768 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/lvs_$n.csv')) if len(r)>10 and r[0].isdigit()]
print('$n', [(r[4].split('::')[-1][6:18], r[7], int(r[-1])/1e6) for r in rows])
PY
done
