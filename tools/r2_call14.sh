#!/bin/bash
cd "$(dirname "$0")/.."
for l in libldpc_b200.so libldpc_RCPPAIR.so; do
  LDPC_LIB_NAME=$l timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c14_$l.json 2> gpurun_out/r2c14_$l.err
  python -c "import json; d=json.load(open('gpurun_out/r2c14_$l.json')); print('$l', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
LDPC_LIB_NAME=libldpc_RCPPAIR.so timeout 600 python tools/parity_fast.py --out gpurun_out/r2c14_parity_rcppair.json --trace-frames 0 --regimes bench,fix_2.0dB,r083_3.5dB,fix_6dB --bench-frames 32768 --frames 8192 2>&1 | cut -c1-420 | tail -5
