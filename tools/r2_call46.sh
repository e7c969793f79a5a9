#!/bin/bash
# one-frame gather kernel as the early-termination default (+ eager syndrome): full GPU suite, ET probe, bench
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c46_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c46_pytest.log
timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 1.5 2.0 3.0 4.0 2> gpurun_out/r2c46_mc_et.err | tee gpurun_out/r2c46_mc_et.jsonl | cut -c1-200; tail -2 gpurun_out/r2c46_mc_et.err
timeout 600 python bench.py > gpurun_out/r2c46_bench.json 2> gpurun_out/r2c46_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c46_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c46_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'mc', round(d['mc']['value'],3), 'mc_et', round(d['mc_early_termination']['value'],2), round(d['mc_early_termination']['frames_per_s']/1e6,2), 'traffic', d['roofline']['traffic'])"
