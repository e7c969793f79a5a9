#!/bin/bash
cd "$(dirname "$0")/.."
LDPC_LIB_NAME=libldpc_ENC2.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c37_launches.csv python tools/mc_et_probe.py wimax_2304_0.5 4.0 > gpurun_out/r2c37.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c37_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:]:
    print(r[ki][:60].replace('\n',' '), r[vi], r[ui])
PY
