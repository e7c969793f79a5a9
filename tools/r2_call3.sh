#!/bin/bash
cd "$(dirname "$0")/.."
LDPC_TRACE_LAUNCH=1 timeout 200 python bench.py --steps 5 --cpu-frames 256 2>&1 >/dev/null | grep "\[ldpc\]" | sort | uniq -c | head -5
for f in 1 2 3; do
  LDPC_PAIR_PER_SM=$f timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c3_bench_tm$f.json 2> gpurun_out/r2c3_bench_tm$f.err
  python -c "import json; d=json.load(open('gpurun_out/r2c3_bench_tm$f.json')); print('tmem per_sm=$f', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4))"
done
