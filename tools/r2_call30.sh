#!/bin/bash
# explicit 32-bit shared-window addressing in the NVRTC builds: GPU suite, JIT gather vs one-frame throughput, default bench
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c30_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c30_pytest.log
timeout 600 python tools/jit_gather_bench.py 48 92 > gpurun_out/r2c30_jit_gather.jsonl 2> gpurun_out/r2c30_jit_gather.err; echo "jit bench rc=$?"; cat gpurun_out/r2c30_jit_gather.jsonl; tail -3 gpurun_out/r2c30_jit_gather.err
timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2c30_bench.json')); print('default bench', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value'],3), 'mc', round(d['mc']['value'],3), 'traffic', d['roofline']['traffic'])"
