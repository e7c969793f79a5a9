#!/bin/bash
# round-2 final state, one GPU: full GPU suite, smoke, all bench lines, reference arm, launch list + full ncu capture, measured parity
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c33_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c33_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c33_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c33_smoke.log
timeout 600 python bench.py > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c33_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c33_bench.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'e2e16', round(d['e2e_f16_ingest']['value'],3), 'e2e8', round(d['e2e_i8_ingest']['value'],3), 'mc', round(d['mc']['value'],3), 'frac', round(d['roofline']['frac'],4))"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c33_bench_ref.json 2>/dev/null; echo "ref rc=$?"
for w in parity576 std576 std2304; do
  timeout 600 python bench.py --workload $w --steps 5 > gpurun_out/r2c33_bench_$w.json 2> gpurun_out/r2c33_bench_$w.err; echo "$w rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c33_launches.csv python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c33_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qc_gather -s 4 -c 1 -o gpurun_out/r2c33_gather python bench.py --steps 3 --warmup 3 --spin 0 --cpu-frames 256 > gpurun_out/r2c33_ncu.log 2>&1
echo "ncu rc=$?"
timeout 1500 python tools/parity_fast.py --out gpurun_out/r2c33_parity_fast.json > gpurun_out/r2c33_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r2c33_parity.log | cut -c1-300
