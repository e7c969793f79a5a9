#!/bin/bash
# fp64 generic path with the hand-rolled tanh / atanh / division: parity tests, then A/B against the libm build
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_montecarlo.py -x -q -m gpu > gpurun_out/r2c22_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c22_pytest.log
for l in libldpc_b200.so libldpc_LIBM.so; do
  for w in parity576 std576 std2304; do
    LDPC_LIB_NAME=$l timeout 600 python bench.py --workload $w --steps 5 > gpurun_out/r2c22_${l}_$w.json 2> gpurun_out/r2c22_${l}_$w.err
    python -c "
import json; d=json.load(open('gpurun_out/r2c22_${l}_$w.json')); print('$l $w', round(d['value'],4), 'Gbit/s', round(d['ms_per_step'],2), 'ms', 'frac', round(d['roofline']['frac'],3))" || tail -3 gpurun_out/r2c22_${l}_$w.err
  done
done
