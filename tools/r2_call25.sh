#!/bin/bash
# fp64 generic path, coefficients in constant memory + branch-free clips: parity tests, benches, instruction count
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f64 or golden or known or config1 or compaction or irregular or ragged or small_batch or tiny" > gpurun_out/r2c25_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c25_pytest.log
for w in parity576 std576 std2304; do
  timeout 600 python bench.py --workload $w --steps 5 > gpurun_out/r2c25_$w.json 2> gpurun_out/r2c25_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c25_$w.json')); print('$w', round(d['value'],4), 'Gbit/s', round(d['ms_per_step'],2), 'ms', 'frac', round(d['roofline']['frac'],3))" || tail -3 gpurun_out/r2c25_$w.err
done


