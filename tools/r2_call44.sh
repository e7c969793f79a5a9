#!/bin/bash
# round-2 final build, eight GPUs: bench line (with the early-termination block) and BASELINE configs[4], the waterfall sweep to BER < 1e-7
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c44_bench_n8.json 2> gpurun_out/r2c44_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r2c44_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r2c44_bench_n8.json'))
print('N=8 value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e8', round(d['e2e_i8_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'mc_et', round(d['mc_early_termination']['value'],2), round(d['mc_early_termination']['frames_per_s']/1e6,1), 'M frames/s')"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/mc_sweep.py \
  --code wimax_2304_0.5 --snr 1.0 2.8 0.2 --fix-odd-check-sign --no-sigma-sq-quirk --max-frames 120000000 --interval-frames 4194304 \
  --min-frame-errors 200 --out gpurun_out/r2c44_mc_sweep_n8_results.json > gpurun_out/r2c44_mc_sweep_n8.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/r2c44_mc_sweep_n8.log | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2c44_mc_sweep_n8_results.json')); print('sweep wall clock', d['wall_clock_seconds'])"
