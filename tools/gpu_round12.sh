#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=30 run tc 900 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu -p no:cacheprovider
TAILN=30 run engine 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -p no:cacheprovider
TAILN=30 run prof_chain 600 python tools/prof_chain.py
TAILN=3 run bench_bf16 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py 5
cat gpurun_out/summary.txt
