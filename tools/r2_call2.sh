#!/bin/bash
# round-2 GPU call: pair kernel tests, A/B bench, ncu capture
cd "$(dirname "$0")/.."
timeout 120 python tools/pair_debug.py > gpurun_out/r2c2_debug.log 2>&1; echo "debug rc=$?"; tail -6 gpurun_out/r2c2_debug.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair or fp32_paths or resident_kernel_equals or specialised_and_table" > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c2_pytest.log
for v in gather scatter regs one; do
  unset LDPC_BENCH_ONE_FRAME LDPC_BENCH_PAIR_REGS LDPC_BENCH_PAIR_SCATTER
  if [ $v = scatter ]; then export LDPC_BENCH_PAIR_SCATTER=1; fi
  if [ $v = one ]; then export LDPC_BENCH_ONE_FRAME=1; fi
  if [ $v = regs ]; then export LDPC_BENCH_PAIR_REGS=1; fi
  timeout 300 python bench.py --steps 10 --cpu-frames 512 > gpurun_out/r2c2_bench_$v.json 2> gpurun_out/r2c2_bench_$v.err
  python -c "import json; d=json.load(open('gpurun_out/r2c2_bench_$v.json')); print('$v', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],3), d['clocks'])"
done
unset LDPC_BENCH_ONE_FRAME LDPC_BENCH_PAIR_REGS LDPC_BENCH_PAIR_SCATTER
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c2_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c2_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2c2_ncu.log
