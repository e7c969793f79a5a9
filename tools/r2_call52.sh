#!/bin/bash
# final build, eight GPUs: bench line
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c52_bench_n8.json 2> gpurun_out/r2c52_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r2c52_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r2c52_bench_n8.json'))
print('N=8 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'e2e8', round(d['e2e_i8_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'allreduce_us', round(d['mc']['allreduce_us'],1), 'mc_et', round(d['mc_early_termination']['value'],2), round(d['mc_early_termination']['frames_per_s']/1e6,1), 'M frames/s')"
