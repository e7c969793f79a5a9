#!/bin/bash
# A/B: one-frame gather kernel (early-termination default) with its previous messages read back from the edge buffer
cd "$(dirname "$0")/.."
for l in libldpc_b200.so libldpc_G1SMEM.so libldpc_b200.so libldpc_G1SMEM.so; do
  echo "== $l"; LDPC_LIB_NAME=$l timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 1.5 2.0 3.0 4.0 2>/dev/null | cut -c1-150
done
LDPC_LIB_NAME=libldpc_G1SMEM.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_montecarlo.py -x -q -m gpu > gpurun_out/r2c51_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c51_pytest.log
