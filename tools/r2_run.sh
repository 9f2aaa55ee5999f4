#!/bin/bash
# Round-2 GPU session helper: runs named steps in separate processes, logs under gpurun_out/.
# usage: tools/r2_run.sh step [step...]   steps: tests smoke overlap bench benchq stages
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
for step in "$@"; do
  case $step in
    tests) TAILN=25 run tests 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rA --tb=short ;;
    testsnew) TAILN=40 run testsnew 1500 python -m pytest tests/test_gpu_atsize.py tests/test_gpu_dropin_reference.py -q -m gpu -p no:cacheprovider -rA --tb=short -s ;;
    smoke) TAILN=2 run smoke 300 python __graft_entry__.py --smoke ;;
    overlap) TAILN=40 run overlap 600 python tools/prof_overlap.py ;;
    overlapc) TAILN=40 run overlapc 600 python tools/prof_overlap.py 4096 64 ;;
    bench) timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_default.json ;;
    benchq) timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== benchq exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_quick.json ;;
    benchref) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== benchref exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_ref.json ;;
    stages) TAILN=12 run stages 300 python tools/prof_stages.py 10 ;;
    *) echo "unknown step $step" ;;
  esac
done
cat gpurun_out/summary.txt
