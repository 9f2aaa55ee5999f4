#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=30 run tc_umma 600 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -p no:cacheprovider -k "umma"
TAILN=30 run tc_query 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu -p no:cacheprovider -k "query"
TAILN=30 run tc_bwd 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu -p no:cacheprovider -k "backward"
TAILN=30 run engine 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -p no:cacheprovider
TAILN=45 run prof_chain 600 python tools/prof_chain.py
TAILN=3 run bench_bf16 900 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
