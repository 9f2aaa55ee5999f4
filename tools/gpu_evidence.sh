#!/bin/bash
# Final-kernel evidence: tests, bench, launch lists (train step, render frame), one ncu --set full capture per hot kernel.
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=3 run tests 1200 python -m pytest tests -q -m gpu -p no:cacheprovider
TAILN=40 run prof_chain 600 python tools/prof_chain.py
TAILN=3 run bench_bf16 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run launches_train 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py 5
run launches_render 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_render.csv python tools/prof_render.py 1
run ncu_wgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_wgrad -s 2 -c 1 -f -o gpurun_out/r01_wgrad python tools/prof_train.py 2
run ncu_dgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_dgrad -s 2 -c 1 -f -o gpurun_out/r01_dgrad python tools/prof_train.py 2
run ncu_fwdtrain 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 3 -c 1 -f -o gpurun_out/r01_fwdtrain python tools/prof_train.py 2
run ncu_fwdinf 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 2 -c 1 -f -o gpurun_out/r01_fwdinf python tools/prof_fwd.py 4
cat gpurun_out/summary.txt
