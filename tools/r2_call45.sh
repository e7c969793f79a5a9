#!/bin/bash
cd "$(dirname "$0")/.."
for v in "" one_gather pair_gather; do
  echo "== ET kernel: ${v:-default}"; LDPC_ET_KERNEL=$v timeout 300 python tools/mc_et_probe.py wimax_2304_0.5 1.5 2.0 3.0 4.0 2>/dev/null | cut -c1-200
done
