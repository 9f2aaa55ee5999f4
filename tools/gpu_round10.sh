#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=40 run micro 600 python tools/prof_chain.py --micro
run ncu_fwd3 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_fwd_v3 python tools/prof_fwd.py 4
cat gpurun_out/summary.txt
