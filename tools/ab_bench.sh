#!/bin/bash
# A/B of library builds on one box: tools/ab_bench.sh libldpc_b200.so libldpc_variant.so ...
# (variants are built with LDPC_LIB_NAME=libldpc_variant.so LDPC_NVCC_EXTRA="-D..." python ldpc-simulator_b200/build_native.py --force)
for l in "$@"; do
  LDPC_BENCH_TRACE=1 LDPC_LIB_NAME=$l python bench.py --cpu-frames 512 2>/tmp/ab_err.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$l', round(d['value'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), d['clocks'])"
  grep "timed steps" /tmp/ab_err.txt | cut -c1-200
  grep warm-up /tmp/ab_err.txt | awk '{print $6}' | sort -n | awk '{a[NR]=$1} END {print "  warm-up steps", NR, "min", a[1], "median", a[int(NR/2)], "max", a[NR]}'
done
