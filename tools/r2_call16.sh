#!/bin/bash
# A/B: sign by multiplication (+ undefined TMEM padding) vs default; pipe-overlap probe; ncu full capture of the default gather kernel
cd "$(dirname "$0")/.."
./tools/pipe_probe3 > gpurun_out/r2c16_pipe_probe3.txt 2>&1; cat gpurun_out/r2c16_pipe_probe3.txt
for l in libldpc_b200.so libldpc_SIGNMUL.so libldpc_PADANY.so libldpc_b200.so; do
  LDPC_LIB_NAME=$l timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c16_$l.json 2> gpurun_out/r2c16_$l.err
  python -c "import json; d=json.load(open('gpurun_out/r2c16_$l.json')); print('$l', round(d['value'],3), 'Gbit/s kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c16_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c16_ncu.log 2>&1
echo "ncu rc=$?"
