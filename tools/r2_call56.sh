#!/bin/bash
# two GPUs, final build: the multi-GPU tests (2-rank NCCL sweep, determinism, one handle per device) and the bench line at N=2
cd "$(dirname "$0")/.."
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_montecarlo.py tests/test_gpu_parity.py -x -q -m gpu -k "two_rank or deterministic or bound_to_its_device or two_host_threads" > gpurun_out/r2c56_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c56_pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c56_bench_n2.json 2> gpurun_out/r2c56_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r2c56_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2c56_bench_n2.json'))
print('N=2 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'allreduce_us', d['mc']['allreduce_us'], 'mc_et', round(d['mc_early_termination']['value'],2))"
