#!/bin/bash
# A/B: row-by-row source order in the one-frame scatter kernel (early termination on wide-row codes, run-time compiled codes)
cd "$(dirname "$0")/.."
for l in libldpc_b200.so libldpc_ROWWISE.so libldpc_b200.so libldpc_ROWWISE.so; do
  echo "== $l"; LDPC_ET_KERNEL=one_frame LDPC_LIB_NAME=$l timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 2.0 4.0 2>/dev/null | cut -c1-150
done
