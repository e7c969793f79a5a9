#!/bin/bash
# eager syndrome check in the first passes of the early-termination kernel: full GPU suite, ET probe, default bench
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c41_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c41_pytest.log
timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 1.5 2.0 3.0 4.0 2> gpurun_out/r2c41_mc_et.err | tee gpurun_out/r2c41_mc_et.jsonl | cut -c1-230; tail -2 gpurun_out/r2c41_mc_et.err
timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c41_bench.json 2> gpurun_out/r2c41_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2c41_bench.json')); print('default bench', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value'],3), 'mc', round(d['mc']['value'],3), 'traffic', d['roofline']['traffic'])"
