#!/bin/bash
# final build of round 2 (handles bound to their device): GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c57_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c57_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c57_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c57_smoke.log
timeout 600 python bench.py > gpurun_out/r2c57_bench.json 2> gpurun_out/r2c57_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c57_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c57_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'mc_et', round(d['mc_early_termination']['value'],2), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])"
