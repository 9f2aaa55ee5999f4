#!/bin/bash
# compute-sanitizer over the GPU tests (small shapes): memcheck on the parity / engine / round-2 entry tests, racecheck on
# the shared-memory kernels' golden tests.  Summaries -> gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out
out=gpurun_out/r02_sanitizer.txt
: > $out
run() { local tool=$1; shift; echo "== compute-sanitizer --tool $tool python -m pytest $*" >> $out; timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitize_$tool.log | sort -u >> $out; }
run memcheck tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_gpu_round2_entries.py tests/test_gpu_tensorcore.py -k "not full_size and not frame_sized and not idx_big and not at_scale"
run racecheck tests/test_gpu_parity.py -k "golden or bit_exact and not big"
cat $out
