#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; local to=$2; shift 2; timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "== $name exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run parity 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider
run engine 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider
run bench_fp32 900 python bench.py --precision fp32 --steps 3 --warmup 3
run ncu_fwd 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_fwd python tools/prof_fwd.py 4
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_render.csv python -c "
import torch, sys
sys.path.insert(0,'.')
import bench, torch_nerf_b200 as tn
from torch_nerf_b200.engine import HotPathEngine
c=tn.NeRF(63,27,precision='bf16').cuda(); f=tn.NeRF(63,27,precision='bf16').cuda()
e=HotPathEngine(c,f,64,128,'bf16')
cam=tn.PerspectiveCamera({'f_x':bench.blender_focal(800),'f_y':bench.blender_focal(800),'img_width':800,'img_height':800}, bench.pose_spherical(30.,-30.,4.), 2.0, 6.0)
e.render_frame(cam); torch.cuda.synchronize(); print('done')
"
cat gpurun_out/summary.txt
