#!/bin/bash
# final state after the source-order change of the check-node phase: GPU suite, smoke, bench, ncu capture, run-time compiled kernels
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c55_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c55_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c55_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c55_smoke.log
timeout 600 python bench.py > gpurun_out/r2c55_bench.json 2> gpurun_out/r2c55_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c55_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c55_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'mc_et', round(d['mc_early_termination']['value'],2), 'frac', round(d['roofline']['frac'],4))"
timeout 600 python tools/jit_gather_bench.py 48 92 > gpurun_out/r2c55_jit_gather.jsonl 2> gpurun_out/r2c55_jit_gather.err; cat gpurun_out/r2c55_jit_gather.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c55_launches.csv python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c55_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c55_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c55_ncu.log 2>&1
echo "ncu rc=$?"
