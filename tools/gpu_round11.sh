#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=40 run configs 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -p no:cacheprovider -k "config"
TAILN=4 run bench1 900 python bench.py --steps 20 --warmup 3
TAILN=4 run bench2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3
TAILN=4 run bench_ref 900 python bench.py --impl reference --steps 3 --warmup 1
cat gpurun_out/summary.txt
