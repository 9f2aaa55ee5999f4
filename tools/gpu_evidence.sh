#!/bin/bash
# Final-kernel evidence: tests, smoke, bench, launch lists (train step, render frame), ncu --set full captures.
# usage: tools/gpu_evidence.sh [mlp]   ("mlp" also re-captures the four tensor-core kernels)
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
TAILN=3 run tests 1200 python -m pytest tests -q -m gpu -p no:cacheprovider
TAILN=2 run smoke 300 python __graft_entry__.py --smoke
TAILN=12 run stages 300 python tools/prof_stages.py 10
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
run launches_train 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py 5
run launches_render 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_render.csv python tools/prof_render.py 1
for k in sample_fine composite_fwd composite_bwd sample_coarse; do
  run ncu_$k 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r01_$k python tools/prof_stages.py 1
done
if [ "$1" = "mlp" ]; then
  run ncu_wgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_wgrad -s 2 -c 1 -f -o gpurun_out/r01_wgrad python tools/prof_train.py 2
  run ncu_dgrad 600 ncu --set full --clock-control none --import-source on -k regex:mlp_dgrad -s 2 -c 1 -f -o gpurun_out/r01_dgrad python tools/prof_train.py 2
  run ncu_fwdtrain 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 3 -c 1 -f -o gpurun_out/r01_fwdtrain python tools/prof_train.py 2
  run ncu_fwdinf 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 2 -c 1 -f -o gpurun_out/r01_fwdinf python tools/prof_fwd.py 4
fi
cat gpurun_out/summary.txt
