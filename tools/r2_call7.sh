#!/bin/bash
# round-2: full GPU suite, bench (all workloads), launch list + ncu full capture of the headline kernel
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c7_pytest.log
timeout 600 python bench.py > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c7_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c7_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'mc', round(d['mc']['value'],3), d['mc']['kernel_ms'], d['mc']['kernel_ms_max'], d['mc']['allreduce_us'], d['mc']['host_sync_us'], 'frac', round(d['roofline']['frac'],4))"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c7_bench_ref.json 2>/dev/null; echo "ref rc=$?"
for w in parity576 std576 std2304; do
  timeout 600 python bench.py --workload $w --steps 5 > gpurun_out/r2c7_bench_$w.json 2> gpurun_out/r2c7_bench_$w.err; echo "$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2c7_bench_$w.json')); print('$w', round(d['value'],4), 'Gbit/s', round(d['ms_per_step'],2), 'ms', 'frac', round(d['roofline']['frac'],3))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c7_launches.csv python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c7_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c7_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c7_ncu.log 2>&1
echo "ncu rc=$?"
