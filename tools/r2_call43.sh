#!/bin/bash
# pool of host staging pipelines: full GPU suite (incl. the two-thread test) + default bench
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c43_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c43_pytest.log
timeout 600 python bench.py > gpurun_out/r2c43_bench.json 2> gpurun_out/r2c43_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c43_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c43_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'mc_et', round(d['mc_early_termination']['value'],2), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])"
