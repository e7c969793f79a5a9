#!/bin/bash
cd "$(dirname "$0")/.."
LDPC_TRACE_LAUNCH=1 timeout 200 python bench.py --steps 5 --cpu-frames 256 2>&1 >/dev/null | grep "\[ldpc\]" | sort | uniq -c | head -3
timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2c4_bench.json')); print('tmem default', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_pair -s 4 -c 1 -o gpurun_out/r2c4_pair_tmem python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c4_ncu.log 2>&1
echo "ncu rc=$?"
