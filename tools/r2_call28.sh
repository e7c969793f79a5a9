#!/bin/bash
# JIT gather kernel: full GPU suite (incl. the new bit-identity test on unregistered base matrices) + throughput on scaled WiMAX r1/2 codes
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c28_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c28_pytest.log
timeout 600 python tools/jit_gather_bench.py 48 92 > gpurun_out/r2c28_jit_gather.jsonl 2> gpurun_out/r2c28_jit_gather.err; echo "jit bench rc=$?"; cat gpurun_out/r2c28_jit_gather.jsonl; tail -3 gpurun_out/r2c28_jit_gather.err
