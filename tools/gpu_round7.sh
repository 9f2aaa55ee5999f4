#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=25 run all_gpu_tests 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
TAILN=75 run prof_chain 600 python tools/prof_chain.py
run ncu_wgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_wgrad -s 2 -c 1 -f -o gpurun_out/prof_wgrad python tools/prof_train.py 2
run ncu_dgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_dgrad -s 2 -c 1 -f -o gpurun_out/prof_dgrad python tools/prof_train.py 2
run ncu_fwdtrain 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 3 -c 1 -f -o gpurun_out/prof_fwdtrain python tools/prof_train.py 2
cat gpurun_out/summary.txt
