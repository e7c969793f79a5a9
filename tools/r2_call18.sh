#!/bin/bash
# A/B: four-team gather schedule (24 warps per SM, 80 registers) vs default
cd "$(dirname "$0")/.."
for l in libldpc_b200.so libldpc_G4.so libldpc_b200.so libldpc_G4.so; do
  LDPC_LIB_NAME=$l timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c18_$l.json 2> gpurun_out/r2c18_$l.err
  python -c "import json; d=json.load(open('gpurun_out/r2c18_$l.json')); print('$l', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'mc', round(d['mc']['value'],3))" || tail -3 gpurun_out/r2c18_$l.err
done
LDPC_LIB_NAME=libldpc_G4.so timeout 600 python tools/parity_fast.py --out gpurun_out/r2c18_parity_g4.json --trace-frames 0 --regimes bench,fix_2.0dB --bench-frames 8192 --frames 4096 2>&1 | cut -c1-300 | tail -3
