#!/bin/bash
# bench line with the early-termination Monte-Carlo block; fresh ncu capture of the headline kernel (traffic record)
cd "$(dirname "$0")/.."
timeout 600 python bench.py > gpurun_out/r2c42_bench.json 2> gpurun_out/r2c42_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c42_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c42_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'mc', round(d['mc']['value'],3), 'frac', round(d['roofline']['frac'],4)); print(json.dumps(d['mc_early_termination'])[:600])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c42_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c42_ncu.log 2>&1
echo "ncu rc=$?"
