#!/bin/bash
# round-2: compute-sanitizer memcheck over every kernel family (legacy part, then the pair / TMEM kernels)
cd "$(dirname "$0")/.."
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_legacy.log python tools/sanitize_small.py legacy > gpurun_out/r2_sanitize_legacy.out 2>&1; echo "memcheck legacy rc=$?"; tail -2 gpurun_out/r2_sanitize_legacy.out; tail -3 gpurun_out/r2_sanitizer_memcheck_legacy.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_pair.log python tools/sanitize_small.py pair > gpurun_out/r2_sanitize_pair.out 2>&1; echo "memcheck pair rc=$?"; tail -2 gpurun_out/r2_sanitize_pair.out; tail -3 gpurun_out/r2_sanitizer_memcheck_pair.log
