#!/bin/bash
cd "$(dirname "$0")/.."
PAIR_DEBUG_VARIANT=one_gather_kernel timeout 120 python tools/pair_debug.py > gpurun_out/r2c11_debug.log 2>&1; echo "debug rc=$?"; tail -5 gpurun_out/r2c11_debug.log | cut -c1-200
for v in gather one_gather; do
  unset LDPC_BENCH_ONE_GATHER
  if [ $v = one_gather ]; then export LDPC_BENCH_ONE_GATHER=1; fi
  LDPC_TRACE_LAUNCH=1 timeout 300 python bench.py --steps 10 --cpu-frames 512 > gpurun_out/r2c11_bench_$v.json 2> gpurun_out/r2c11_bench_$v.err
  grep "\[ldpc\]" gpurun_out/r2c11_bench_$v.err | sort | uniq -c | head -2
  python -c "import json; d=json.load(open('gpurun_out/r2c11_bench_$v.json')); print('$v', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],3), 'i8', round(d['e2e_i8_ingest']['value'],3), d['e2e_i8_ingest']['decision_bit_agreement_vs_fp32_ingest'])"
done
