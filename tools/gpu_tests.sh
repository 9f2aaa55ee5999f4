#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context) and keeps logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; local to=$2; shift 2
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "== $name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run parity 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider
run umma_k 300 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -k umma_kmajor -p no:cacheprovider
run umma_mn 300 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -k umma_mnmajor -p no:cacheprovider
run bf16 300 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -k bf16 -p no:cacheprovider
run smoke 300 python __graft_entry__.py --smoke
cat gpurun_out/summary.txt
