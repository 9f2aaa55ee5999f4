#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run tc_fwd 600 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -k "umma or query" -p no:cacheprovider
run tc_bwd 600 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -k "backward" -p no:cacheprovider -x
run engine 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider
run parity 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider
TAILN=3 run bench_bf16 900 python bench.py --precision bf16 --steps 10 --warmup 3
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 300 --csv --log-file gpurun_out/launches_train.csv python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
