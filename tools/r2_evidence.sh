#!/bin/bash
# Round-2 evidence on one B200: launch lists (training step, 800x800 frame), ncu --set full captures of the hot kernels and
# of the kernels that changed this round, condensed to text on the box by tools/ncu_summary.py.  Outputs: gpurun_out/r02_*
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-3} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run launches_train 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/r02_launches_train_step.csv python tools/prof_train.py 5
run launches_render 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r02_launches_render_frame.csv python tools/prof_render.py 1
cap() { # name, kernel regex, skip, script, args...
  local name=$1 rx=$2 skip=$3; shift 3
  run ncu_$name 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r02_$name python "$@"
  python tools/ncu_summary.py gpurun_out/r02_$name.ncu-rep gpurun_out/r02_ncu_$name.txt > /dev/null 2>&1
  rm -f gpurun_out/r02_$name.ncu-rep
}
cap wgrad mlp_wgrad 2 tools/prof_train.py 2
cap dgrad mlp_dgrad 2 tools/prof_train.py 2
cap fwd_train mlp_fwd 3 tools/prof_train.py 2
cap fwd_inference mlp_fwd_kernel 2 tools/prof_fwd.py 4
cap composite_bwd_mse composite_bwd 2 tools/prof_train.py 2
cap train_prologue train_prologue 1 tools/prof_train.py 2
cap sample_fine sample_fine 1 tools/prof_stages.py 1
TAILN=12 run stages 300 python tools/prof_stages.py 10
cp gpurun_out/stages.log gpurun_out/r02_stage_kernels.txt
cat gpurun_out/summary.txt
