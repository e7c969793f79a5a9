#!/bin/bash
# generic fp64 path: launch list of a parity576 step + full ncu capture of one check-node launch
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c23_launches_parity576.csv python bench.py --workload parity576 --steps 2 --warmup 3 --cpu-frames 64 > gpurun_out/r2c23_list.log 2>&1; echo "list rc=$?"
python tools/launch_list_summary.py gpurun_out/r2c23_launches_parity576.csv 2>/dev/null | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_check_nodes -s 30 -c 1 -o gpurun_out/r2c23_check_nodes python bench.py --workload parity576 --steps 2 --warmup 3 --cpu-frames 64 > gpurun_out/r2c23_ncu.log 2>&1; echo "ncu rc=$?"
