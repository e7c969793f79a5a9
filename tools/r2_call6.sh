#!/bin/bash
# round-2: full GPU test suite + smoke + bench (default, reference arm) + parity measurement
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c6_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c6_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c6_smoke.log
timeout 600 python bench.py > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c6_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c6_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'agree16', d['e2e_f16_ingest']['decision_bit_agreement_vs_fp32_ingest'], 'mc', round(d['mc']['value'],3), d['mc']['allreduce_us'], d['mc']['host_sync_us'], 'frac', round(d['roofline']['frac'],4), 'smem', round(d['roofline']['smem']['frac'],3))"
timeout 600 python tools/parity_fast.py --out gpurun_out/r2_parity_fast.json --trace-frames 256 > gpurun_out/r2c6_parity.log 2>&1; echo "parity rc=$?"
