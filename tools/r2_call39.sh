#!/bin/bash
# tiled device encoder: Monte-Carlo tests on the variant, then the early-termination probe A/B
cd "$(dirname "$0")/.."
LDPC_LIB_NAME=libldpc_ENC5.so timeout 900 python -m pytest tests/test_gpu_montecarlo.py -x -q -m gpu > gpurun_out/r2c39_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c39_pytest.log
for l in libldpc_b200.so libldpc_ENC5.so; do
  echo "== $l"; LDPC_LIB_NAME=$l timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 2.0 3.0 4.0 2> gpurun_out/r2c39_$l.err | tee gpurun_out/r2c39_mc_et_$l.jsonl | cut -c1-230; tail -2 gpurun_out/r2c39_$l.err
done
