#!/bin/bash
# round-2, two GPUs: 2-rank NCCL Monte-Carlo test, bench at N=2, plain run of the sanitizer script
cd "$(dirname "$0")/.."
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_montecarlo.py tests/test_gpu_parity.py -x -q -m gpu -k "two_rank or deterministic" > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c9_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r2c9_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2c9_bench_n2.json'))
print('N=2 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['e2e']['per_rank_gbit_s'], 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'allreduce_us', d['mc']['allreduce_us'], 'host_us', d['mc']['host_sync_us'])"
(timeout 600 python tools/sanitize_small.py legacy; timeout 600 python tools/sanitize_small.py pair) > gpurun_out/r2_sanitize_small_plain.log 2>&1; echo "plain sanitize rc=$?"; tail -2 gpurun_out/r2_sanitize_small_plain.log | cut -c1-300
