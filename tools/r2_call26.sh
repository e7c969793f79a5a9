#!/bin/bash
# e2e (host buffers in, results out) against the chunk size of the host pipeline
cd "$(dirname "$0")/.."
for mb in 64 32 16 8 4; do
  LDPC_HOST_CHUNK_MB=$mb timeout 300 python bench.py --steps 10 --cpu-frames 256 > gpurun_out/r2c26_chunk$mb.json 2> gpurun_out/r2c26_chunk$mb.err
  python -c "import json; d=json.load(open('gpurun_out/r2c26_chunk$mb.json')); print('chunk $mb MB: e2e', round(d['e2e']['value'],3), 'h2d GB/s', round(d['e2e']['h2d_gbs_whole_job'],2), 'f16', round(d['e2e_f16_ingest']['value'],3), 'i8', round(d['e2e_i8_ingest']['value'],3), 'value', round(d['value'],3))" || tail -3 gpurun_out/r2c26_chunk$mb.err
done
