#!/bin/bash
# round-2, eight GPUs: BASELINE configs[4] -- waterfall sweep to BER < 1e-7 with in-kernel Philox and NCCL counter all-reduce
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/mc_sweep.py \
  --code wimax_2304_0.5 --snr 1.0 2.8 0.2 --fix-odd-check-sign --no-sigma-sq-quirk --max-frames 120000000 --interval-frames 4194304 \
  --min-frame-errors 200 --out gpurun_out/r2_mc_sweep_n8_results.json > gpurun_out/r2_mc_sweep_n8.log 2>&1; echo "sweep rc=$?"; tail -4 gpurun_out/r2_mc_sweep_n8.log | cut -c1-400
