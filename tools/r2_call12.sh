#!/bin/bash
cd "$(dirname "$0")/.."
for l in libldpc_b200.so libldpc_NOSMEM.so libldpc_NOLG2.so libldpc_NOEX2.so; do
  LDPC_LIB_NAME=$l timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c12_$l.json 2> gpurun_out/r2c12_$l.err
  python -c "import json; d=json.load(open('gpurun_out/r2c12_$l.json')); print('$l', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
