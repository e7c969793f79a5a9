#!/bin/bash
# round-2, eight GPUs: host->device ceiling, bench at N=8 and N=4
cd "$(dirname "$0")/.."
nvidia-smi -L | wc -l; nproc; numactl -H 2>/dev/null | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/h2d_probe.py > gpurun_out/r2_h2d_probe_n8.json 2> gpurun_out/r2_h2d_probe_n8.err; echo "h2d rc=$?"; cat gpurun_out/r2_h2d_probe_n8.json | cut -c1-700
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2c10_bench_n$n.json 2> gpurun_out/r2c10_bench_n$n.err; echo "bench n$n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c10_bench_n$n.json'))
print('N=$n value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), [round(x,2) for x in d['e2e']['per_rank_gbit_s']], 'h2d', round(d['e2e']['h2d_gbs_whole_job'],1), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'h2d16', round(d['e2e_f16_ingest']['h2d_gbs_whole_job'],1), 'mc', round(d['mc']['value'],3), 'allreduce_us', d['mc']['allreduce_us'], 'host_us', d['mc']['host_sync_us'])"
done
