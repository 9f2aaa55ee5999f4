#!/bin/bash
# Short end-of-session check on a B200 box: all GPU tests, smoke, stage timings, the default bench line, the two launch
# lists and one ncu capture of the fine sampler (a subset of tools/gpu_evidence.sh).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 200 python tools/prof_stages.py 10 | tee gpurun_out/stages.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench $?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py 5 > gpurun_out/launches_train.log 2>&1; echo lt $?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_render.csv python tools/prof_render.py 1 > gpurun_out/launches_render.log 2>&1; echo lr $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sample_fine -s 1 -c 1 -f -o gpurun_out/r01_sample_fine python tools/prof_stages.py 1 > gpurun_out/ncu_sample_fine.log 2>&1; echo ncu $?
