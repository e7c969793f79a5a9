#!/bin/bash
# round-2 final state, N GPUs: 2-rank NCCL tests (N >= 2) and the bench line at N
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi -L | wc -l
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_montecarlo.py tests/test_gpu_parity.py -x -q -m gpu -k "two_rank or deterministic" > gpurun_out/r2c34_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c34_pytest_n2.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c34_bench_n$N.json 2> gpurun_out/r2c34_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r2c34_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2c34_bench_n$N.json'))
print('N=$N value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), [round(x,2) for x in d['e2e']['per_rank_gbit_s']], 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'e2e8', round(d['e2e_i8_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'allreduce_us', round(d['mc']['allreduce_us'],1), 'host_us', round(d['mc']['host_sync_us'],1))"
