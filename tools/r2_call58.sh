#!/bin/bash
# final build, N GPUs (argument): bench line
cd "$(dirname "$0")/.."
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c58_bench_n$N.json 2> gpurun_out/r2c58_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r2c58_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2c58_bench_n$N.json'))
print('N=$N value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'e2e8', round(d['e2e_i8_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'allreduce_us', round(d['mc']['allreduce_us'],1), 'mc_et', round(d['mc_early_termination']['value'],2), round(d['mc_early_termination']['frames_per_s']/1e6,1), 'M frames/s')"
