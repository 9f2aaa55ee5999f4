#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-ray rendering hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp32]

Metric: train rays/s, one training step = coarse (64) + fine (64+128) forward and backward of a 4096-ray batch
per GPU on a lego-shaped synthetic scene (800x800 cameras, near/far 2/6), gradients all-reduced over NCCL, Adam
step included (BASELINE.json configs[1]; weak scaling for N > 1 = configs[3] shape).  The same line also reports
the 800x800 full-frame render latency (configs[2]) under "render".

`--impl reference` times the reference's CPU path (the numpy oracle port, all host threads) on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SC, SF = 64, 128
IMG = 800
FLOP_FWD_PER_EVAL = 2 * 593408
FLOP_TRAIN_PER_EVAL = 3489024
METRIC = "train rays/s (64+128 samples, fwd+bwd); 800x800 render ms @1/2/4/8 B200"


def recorded_traffic(kernel, rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on `rows` rows, as captured with
    `ncu --set full` and recorded in profiles/traffic.json; None when no capture of exactly this shape exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    for rec in json.load(open(path)).get(kernel, []):
        if rec.get("rows") == rows:
            return rec.get("dram_bytes")
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ synthetic scene
def pose_spherical(theta_deg, phi_deg, radius):
    """Blender-shaped camera-to-world (reference utils/data/load_blender.py:78-109)."""
    import torch

    t = torch.eye(4)
    t[2, 3] = radius
    phi, th = phi_deg / 180.0 * math.pi, theta_deg / 180.0 * math.pi
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(phi), -math.sin(phi), 0], [0, math.sin(phi), math.cos(phi), 0], [0, 0, 0, 1.0]])
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1.0]])
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return flip @ (rt @ (rp @ t))


def blender_focal(w, angle=0.6911112070083618):
    return 0.5 * w / math.tan(0.5 * angle)


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML every 10 ms when pynvml is importable
    (nvidia_ml_py), else one nvidia-smi query every 200 ms."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx, self.samples, self._stop, self._th = gpu_index, [], threading.Event(), None
        self.source = "nvidia-smi"

    def _run_nvml(self):
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while not self._stop.is_set():
            mask = int(reasons_fn(h))
            row = [str(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))), str(mx)]
            row += ["Active" if mask & self.BITS[n] else "Not Active" for n in self.NAMES]
            self.samples.append(row)
            self._stop.wait(0.01)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples), "source": self.source}


# ------------------------------------------------------------------------------------------------ reference arms
def cpu_train_step_rays_per_s(rays, reps, seed=0):
    """Fallback CPU baseline when baseline/_ref is not staged: the oracle's training iteration (train.py:130-218
    restated in numpy) on the host cores."""
    from oracle import nerf_oracle as orc

    rng = np.random.default_rng(seed)
    pc, pf = orc.init_nerf_params(seed=1), orc.init_nerf_params(seed=2)
    focal = orc.blender_focal(IMG)
    c2w = orc.pose_spherical(30.0, -30.0, 4.0)
    coords = orc.screen_coords(IMG, IMG)[rng.choice(IMG * IMG, size=rays, replace=False)]
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, IMG, IMG), c2w, 2.0, IMG, IMG, False)
    target = rng.random((rays, 3), dtype=np.float32)
    times = []
    for _ in range(reps):
        u = [rng.random((rays, k), dtype=np.float32) for k in (SC, SC, SF, SF)]
        t0 = time.perf_counter()
        orc.train_step_grads(pc, pf, o, d, 2.0, 6.0, SC, SF, target, *u)
        times.append(time.perf_counter() - t0)
    return rays / float(np.median(times)), float(np.median(times))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _ref_scene(img):
    import torch

    focal = blender_focal(img)
    intr = {"f_x": focal, "f_y": focal, "img_width": img, "img_height": img}
    return intr, pose_spherical(30.0, -30.0, 4.0)


def ref_train_rays_per_s(device, rays, steps, warmup, threads=None):
    """The UNMODIFIED reference (baseline/_ref via baseline/ref_harness.py) running train.py:130-218 on `device`:
    median seconds per step over `steps` steps of `rays` rays of an 800x800 camera.  CPU: wall clock; CUDA: events."""
    import torch

    from baseline import ref_harness as rh

    if threads is not None:
        torch.set_num_threads(threads)
    sess = rh.RefSession(device)
    intr, c2w = _ref_scene(IMG)
    gt = torch.rand(IMG * IMG, 3)
    cuda = torch.device(device).type == "cuda"
    times = []
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sess.train_step(gt, intr, c2w, 2.0, 6.0, rays)
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            t0 = time.perf_counter()
            sess.train_step(gt, intr, c2w, 2.0, 6.0, rays)
            dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = float(np.median(times))
    return rays / med, med


def ref_render_ms(device, img, reps, warmup, threads=None):
    """The unmodified reference running render.py:58-107 (coarse + fine, num_ray_batch = H*W // 4096) on `device`."""
    import torch

    from baseline import ref_harness as rh

    if threads is not None:
        torch.set_num_threads(threads)
    sess = rh.RefSession(device)
    intr, c2w = _ref_scene(img)
    cuda = torch.device(device).type == "cuda"
    times = []
    for i in range(warmup + reps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = sess.render_frame(intr, c2w, 2.0, 6.0, (img, img))
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    del out
    return float(np.median(times)) * 1e3


def cpu_baseline_block(quick):
    """cpu_baseline of the bench line: the reference's own CPU path on the box's host cores, bounded (10-30 s)."""
    from baseline import ref_harness as rh

    cores = host_cores()
    ok, why = rh.available()
    if ok:
        rays, steps = 1024, (2 if quick else 3)
        rps, med = ref_train_rays_per_s("cpu", rays, steps, 1, threads=cores)
        return {"value": rps, "unit": "rays/s", "cores": cores, "kind": "reference", "ms_per_step": med * 1e3,
                "sample": f"{steps} x {rays}-ray training steps (+1 warm-up) of the unmodified reference modules "
                          f"(baseline/_ref: VolumeRenderer.render_scene coarse+fine, MSE, backward, Adam) on CPU, "
                          f"torch.set_num_threads({cores}), median; rays/s scales linearly to the 4096-ray batch"}
    rps, med = cpu_train_step_rays_per_s(256, 3)
    return {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "ms_per_step": med * 1e3,
            "sample": f"3 x 256-ray training iterations of the oracle (numpy port; {why}), median"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from baseline import ref_harness as rh

    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: thread counts are set explicitly below
    ok, why = rh.available()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 4))
    extra = {}
    if ok:
        rays = 1024
        rps, med = ref_train_rays_per_s("cpu", rays, steps, max(1, min(args.warmup, 1)), threads=cores)
        kind = "reference"
        sample = (f"{steps} x {rays}-ray training steps of the unmodified reference modules (baseline/_ref) on CPU, "
                  f"torch.set_num_threads({cores}), median")
        if not args.quick:
            # the reference's own single-thread setting (runner_utils.py:427) and config C1
            rps1, med1 = ref_train_rays_per_s("cpu", rays, 1, 0, threads=1)
            c1_all = ref_render_ms("cpu", 100, 2, 0, threads=cores)
            c1_one = ref_render_ms("cpu", 100, 1, 0, threads=1)
            extra = {"train_1_thread": {"value": rps1, "unit": "rays/s", "cores": 1, "ms_per_step": med1 * 1e3,
                                        "sample": f"1 x {rays}-ray training step, torch.set_num_threads(1)"},
                     "c1_render_100x100": {"all_cores_ms": c1_all, "one_thread_ms": c1_one, "cores": cores,
                                           "sample": "100x100 frame, 64 + (64+128) samples, num_ray_batch 2 (render.py:58-107); "
                                                     "median of 2 frames (all cores), 1 frame (1 thread)"}}
            import torch

            torch.set_num_threads(cores)
    else:
        rays = 256
        rps, med = cpu_train_step_rays_per_s(rays, steps)
        kind = "port"
        sample = f"{steps} x {rays}-ray training iterations (oracle numpy port: {why})"
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"train step (coarse 64 + fine 64+128, fwd+bwd, Adam), {rays}-ray sample of the {args.rays}-ray "
                               f"batch, lego-shaped synthetic scene, 800x800 cameras, near/far 2/6, reference CPU path",
                   "rays_per_step": rays},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample, **extra},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


def reference_gpu_block(dev, n_rays):
    """Same-box comparator: the unmodified reference modules on torch-CUDA on this
    B200 -- cuBLAS fp32 GEMMs + ATen elementwise kernels + autograd, exactly what a user of the reference runs today."""
    import torch

    from baseline import ref_harness as rh

    ok, why = rh.available()
    if not ok:
        return {"unavailable": why}
    out = {"impl": "unmodified reference modules (baseline/_ref) on torch-CUDA, fp32 (TF32 off = torch default)",
           "torch": torch.__version__}
    with torch.cuda.device(dev):
        rps, med = ref_train_rays_per_s(dev, n_rays, 5, 2)
        out["train"] = {"rays_per_step": n_rays, "ms_per_step": med * 1e3, "rays_per_s": rps,
                        "sample": "median of 5 steps after 2 warm-ups, CUDA events around train.py:130-218"}
        torch.cuda.empty_cache()
        out["c1_render_100x100_ms"] = ref_render_ms(dev, 100, 3, 1)
        out["render_800x800_ms"] = ref_render_ms(dev, IMG, 1, 1)
        torch.cuda.empty_cache()
    return out


def dropin_block(tn, dev, n_rays, eng_render):
    """The same workloads through the reference-facing plugin API of this package -- VolumeRenderer.render_scene with
    StratifiedSampler / QuadratureIntegrator / PrimitiveCube(NeRF) and torch autograd + torch.optim.Adam, driven exactly
    like train.py:130-218 and render.py:58-107 -- instead of the fused engine: what a maintainer gets by swapping the
    classes in runner_utils.py and nothing else ((N,S,3) sample points are materialised here, as the API demands)."""
    import torch

    out = {}
    intr, c2w = _ref_scene(IMG)
    for precision in ("bf16", "fp32"):
        torch.manual_seed(0)
        nets = [tn.NeRF(63, 27, precision=precision).to(dev) for _ in range(2)]
        enc = {"coord_enc": tn.PositionalEncoder(3, 10, True), "dir_enc": tn.PositionalEncoder(3, 4, True)}
        scenes = [tn.PrimitiveCube(net, enc) for net in nets]
        cam = tn.PerspectiveCamera(intr, c2w, 2.0, 6.0)
        ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), tn.StratifiedSampler(), cam)
        opt = torch.optim.Adam([p for net in nets for p in net.parameters()], lr=5e-4, eps=1e-8)
        sched = torch.optim.lr_scheduler.ExponentialLR(opt, (5e-5 / 5e-4) ** (1.0 / 300000))
        loss_fn = torch.nn.MSELoss()
        gt = torch.rand(IMG * IMG, 3)
        dev_i = dev.index if dev.index is not None else torch.cuda.current_device()

        def step(pix=None):
            opt.zero_grad()
            ren.camera = tn.PerspectiveCamera(intr, c2w, 2.0, 6.0)
            pred_c, idx, w_c = ren.render_scene(scenes[0], n_rays, SC, False, dev_i, pixel_indices=pix)
            loss = loss_fn(gt[idx].to(dev), pred_c)
            pred_f, idx_f, _ = ren.render_scene(scenes[1], n_rays, (SC, SF), False, dev_i, pixel_indices=idx, weights=w_c)
            loss = loss + loss_fn(gt[idx_f].to(dev), pred_f)
            loss.backward()
            opt.step()
            sched.step()

        steps = 10 if precision == "bf16" else 3
        for _ in range(3 if precision == "bf16" else 1):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[f"train_{precision}"] = {"rays_per_step": n_rays, "ms_per_step": ms, "rays_per_s": n_rays / (ms * 1e-3), "steps": steps}
        if precision == "bf16":
            # the same step with the pixel batch GIVEN (train.py:150-169 passes pixel_indices during its warm-up epochs):
            # without the reference's host-side np.random.choice(H*W, n, replace=False) of volume_renderer.py:122-128, which
            # this package mirrors and which alone costs ~6-10 ms of host time per pass at 800x800
            gpix = torch.Generator().manual_seed(5)
            pixs = [torch.randperm(IMG * IMG, generator=gpix)[:n_rays] for _ in range(steps + 2)]
            for i in range(2):
                step(pixs[i])
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                step(pixs[2 + i])
            e1.record()
            torch.cuda.synchronize()
            ms_p = e0.elapsed_time(e1) / steps
            t_c = time.perf_counter()
            for _ in range(3):
                np.random.choice(IMG * IMG, size=[n_rays], replace=False)
            out["train_bf16_pixels_given"] = {"rays_per_step": n_rays, "ms_per_step": ms_p, "rays_per_s": n_rays / (ms_p * 1e-3),
                                              "steps": steps, "host_np_random_choice_ms": (time.perf_counter() - t_c) / 3 * 1e3}
        # config C1: 100x100 frame, num_ray_batch = 10000 // 4096 = 2 (render.py:81-99)
        intr1, c2w1 = _ref_scene(100)
        ren.camera = tn.PerspectiveCamera(intr1, c2w1, 2.0, 6.0)

        def frame():
            with torch.no_grad():
                img, idx, wts = ren.render_scene(scenes[0], 10000, SC, False, dev_i, num_ray_batch=2)
                img, _, _ = ren.render_scene(scenes[1], 10000, (SC, SF), False, dev_i, pixel_indices=idx, weights=wts, num_ray_batch=2)
            return img.clamp(0.0, 1.0)

        frame()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            frame()
        e1.record()
        torch.cuda.synchronize()
        out[f"c1_render_100x100_{precision}_ms"] = e0.elapsed_time(e1) / 5
        del nets, scenes, opt
        torch.cuda.empty_cache()
    # the fused engine on config C1 for comparison
    intr1, c2w1 = _ref_scene(100)
    cam1 = tn.PerspectiveCamera(intr1, c2w1, 2.0, 6.0)
    eng_render.render_frame(cam1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng_render.render_frame(cam1)
    e1.record()
    torch.cuda.synchronize()
    out["c1_render_100x100_engine_bf16_ms"] = e0.elapsed_time(e1) / 5
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import torch_nerf_b200 as tn
    from torch_nerf_b200.engine import HotPathEngine
    from torch_nerf_b200.parallel import allreduce_mean_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # weak scaling (default): args.rays per GPU; --global-rays G (config C4 as written): G / world rays per GPU
    if args.global_rays:
        assert args.global_rays % world == 0, "--global-rays must be divisible by the number of GPUs"
        n_rays = args.global_rays // world
    else:
        n_rays = args.rays
    torch.manual_seed(0)  # identical replicas of both networks on every rank
    coarse = tn.NeRF(63, 27, precision=args.precision).to(dev)
    fine = tn.NeRF(63, 27, precision=args.precision).to(dev)
    eng = HotPathEngine(coarse, fine, SC, SF, precision=args.precision)
    # N > 1: gradient exchange fused with Adam over NVLink peer memory (parallel.PeerExchange); falls back to NCCL
    from torch_nerf_b200.parallel import make_exchange

    px = make_exchange(2 * 595844, dev, world)
    flat = eng.enable_flat_params(px.grad if px is not None else None)
    eng_render = eng if args.precision == "bf16" else HotPathEngine(coarse, fine, SC, SF, precision="bf16")
    # runner_utils.py:690-711: Adam(lr 5e-4, eps 1e-8), ExponentialLR gamma = (5e-5/5e-4)^(1/300000); one Adam over both
    # networks' parameters -- here as ONE flat parameter that all 44 tensors alias (elementwise-identical update)
    from torch_nerf_b200.optim import FlatAdam

    opt = FlatAdam([flat.param], lr=5e-4, eps=1e-8)  # torch.optim.Adam's state and math, one kernel launch (csrc/optim.cu)
    opt.grad_scale = 1.0 / world
    opt.exchange = px
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, (5e-5 / 5e-4) ** (1.0 / 300000))
    torch.manual_seed(1234 + rank)  # per-rank uniform stream / pixels
    focal = blender_focal(IMG)
    poses = [pose_spherical(th, -30.0, 4.0) for th in np.linspace(-180, 180, 9)[:-1]]
    cams = [tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": IMG, "img_height": IMG}, p, 2.0, 6.0) for p in poses]
    total_steps = args.steps + args.warmup
    g = torch.Generator().manual_seed(99 + rank)
    host_pix = [torch.randperm(IMG * IMG, generator=g)[:n_rays].contiguous().pin_memory() for _ in range(total_steps)]
    host_tgt = [torch.rand((n_rays, 3), generator=g).pin_memory() for _ in range(total_steps)]
    dev_pix = [p.to(dev) for p in host_pix]
    dev_tgt = [t.to(dev) for t in host_tgt]
    losses_host = torch.zeros((total_steps, 2)).pin_memory()
    losses_dev = torch.zeros((2,), device=dev)
    in_pix = torch.empty((n_rays,), device=dev, dtype=torch.int64)
    in_tgt = torch.empty((n_rays, 3), device=dev)

    use_graph = args.precision == "bf16" and not args.no_graph

    def step(i, e2e):
        cam = cams[i % len(cams)]
        if use_graph:
            # the iteration replayed from a CUDA graph: inputs go straight from pinned host memory (e2e) or device memory
            # into the graph's static buffers
            losses = eng.train_pixels_graph(cam, host_pix[i] if e2e else dev_pix[i], host_tgt[i] if e2e else dev_tgt[i], False)
        else:
            if e2e:
                in_pix.copy_(host_pix[i], non_blocking=True)
                in_tgt.copy_(host_tgt[i], non_blocking=True)
                pix, tgt = in_pix, in_tgt
            else:
                pix, tgt = dev_pix[i], dev_tgt[i]
            losses = eng.train_pixels(cam, pix, tgt, False, loss_out=losses_dev)
        if px is None:
            allreduce_mean_(flat.grad, world, scale=False)  # one NCCL all-reduce (sum) of the 4.77 MB flat gradient buffer
        opt.step()  # with the peer exchange: sum over the ranks + Adam in ONE launch (csrc/dp_exchange.cu)
        sched.step()
        if e2e:
            losses_host[i].copy_(losses, non_blocking=True)

    def timed(e2e):
        for i in range(args.warmup):
            step(i, e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = eng.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.warmup, total_steps):
            step(i, e2e)
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, eng.launches - l0 + (total_steps - args.warmup)  # + one Adam launch per step

    # ---- N > 1: the sharded step must equal the 1-rank step on the concatenated batch: every
    #      rank takes its slice of ONE global pixel batch and ONE global uniform stream, the gradient buffers are summed
    #      by the exchange that the timed loop uses, and rank 0 recomputes the whole batch alone
    parity = None
    if world > 1 and args.precision == "bf16":
        gcpu = torch.Generator().manual_seed(4242)
        G = n_rays * world
        gpix = torch.randperm(IMG * IMG, generator=gcpu)[:G].to(dev)
        gtgt = torch.rand((G, 3), generator=gcpu).to(dev)
        gdev = torch.Generator(device=dev).manual_seed(4242)
        gu = [torch.rand((G, k), device=dev, generator=gdev) for k in (SC, SC, SF, SF)]
        lo, hi = rank * n_rays, (rank + 1) * n_rays
        eng.train_pixels(cams[0], gpix[lo:hi].contiguous(), gtgt[lo:hi].contiguous(), False, uniforms=[u[lo:hi] for u in gu])
        if px is not None:
            px.allreduce_sum_()
        else:
            allreduce_mean_(flat.grad, world, scale=False)
        summed = flat.grad.clone() / world
        torch.cuda.synchronize()
        if rank == 0:
            eng.train_pixels(cams[0], gpix, gtgt, False, uniforms=gu)
            ref = flat.grad
            scale = float(ref.abs().max())
            parity = {"global_rays": G, "max_abs_diff_over_max": float((summed - ref).abs().max()) / scale,
                      "cosine": float(torch.dot(summed, ref) / (summed.norm() * ref.norm())),
                      "what": "all-reduced gradients of the N sharded steps / N vs rank 0 recomputing the concatenated "
                              "batch (same pixels, targets, and rows of one global uniform stream)"}
        del gu, gpix, gtgt, summed
        eng.release_workspaces()
        torch.cuda.empty_cache()
        dist.barrier()

    with ClockSampler(local) as clk:
        ms_dev, launches = timed(False)
        time.sleep(2.0)  # let the board return to its idle power state so both loops start from the same condition
        ms_e2e, _ = timed(True)
    clocks = clk.summary()
    # ---- config C4 as written: a 32768-ray GLOBAL batch split over the GPUs (strong scaling), 5 steps after 3 warm-ups
    c4 = None
    if args.precision == "bf16" and not args.no_c4 and not args.global_rays and 32768 % world == 0:
        nr4 = 32768 // world
        g4 = torch.Generator().manual_seed(77 + rank)
        pix4 = [torch.randperm(IMG * IMG, generator=g4)[:nr4].to(dev) for _ in range(8)]
        tgt4 = [torch.rand((nr4, 3), generator=g4).to(dev) for _ in range(8)]

        def step4(i):
            eng.train_pixels_graph(cams[i % len(cams)], pix4[i], tgt4[i], False) if use_graph else \
                eng.train_pixels(cams[i % len(cams)], pix4[i], tgt4[i], False, loss_out=losses_dev)
            if px is None:
                allreduce_mean_(flat.grad, world, scale=False)
            opt.step()
            sched.step()

        for i in range(3):
            step4(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3, 8):
            step4(i)
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 5
        if world > 1:
            t4 = torch.tensor([ms4], device=dev)
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
            ms4 = float(t4.item())
        c4 = {"global_rays": 32768, "rays_per_gpu": nr4, "n_gpus": world, "ms_per_step": ms4, "rays_per_s": 32768 / (ms4 * 1e-3),
              "scaling": "strong", "steps": 5}
        del pix4, tgt4
        eng.release_workspaces()
        torch.cuda.empty_cache()
    rays_total = n_rays * world * args.steps
    value = rays_total / (ms_dev * 1e-3)
    e2e_value = rays_total / (ms_e2e * 1e-3)
    pk = peaks()

    # ---- rooflines, measured live with CUDA events on the launching stream (kernels timed alone -> burst peaks):
    #   roofline       = the step's dominant kernel, mlp_wgrad_kernel on the fine pass's rows (HBM-bound by construction)
    #   roofline_mlp   = the tensor-core forward chain on the same rows (tensor-bound)
    roof = None
    roof_mlp = None
    render = None
    if rank == 0:
        time.sleep(3.0)  # the kernels below are timed ALONE against burst peaks: start them from the idle power state
        if True:
            lib = tn._lib.load()
            P = tn._lib.ptr
            m = n_rays * (SC + SF)
            ray_o = torch.randn(n_rays, 3, device=dev)
            ray_d = torch.randn(n_rays, 3, device=dev)
            t = torch.rand(n_rays, SC + SF, device=dev) * 4 + 2
            sig = torch.empty(m, device=dev)
            rgb = torch.empty(m, 3, device=dev)
            packed = fine.packed_weights()

            def timed_kernel(fn, reps=10):
                for _ in range(3):
                    fn()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                ev0.record()
                for _ in range(reps):
                    fn()
                ev1.record()
                torch.cuda.synchronize()
                return ev0.elapsed_time(ev1) * 1e-3 / reps

            def k_fwd():
                tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), SC + SF, m,
                                                        P(sig), P(rgb), None, tn._lib.stream()), "fwd")
            sec = timed_kernel(k_fwd)
            ach = m * FLOP_FWD_PER_EVAL / sec / 1e12
            roof_mlp = {"bound": "tensor", "kernel": f"mlp_fwd_kernel (tcgen05 forward chain, {m} rows, inference form)",
                        "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "traffic": None,
                        "peak_source": pk["src"] + " (burst: kernel timed alone)", "launch_ms": sec * 1e3}
            # wgrad alone: fill cache + scratch with one full forward/backward, then re-run only the wgrad phase
            cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device=dev)
            scratch = torch.empty(lib.nerf_mlp_bf16_bwd_scratch_bytes(m), dtype=torch.uint8, device=dev)
            g_s = torch.randn(m, device=dev) * 1e-3
            g_c = torch.randn(m, 3, device=dev) * 1e-3
            grads = [torch.empty_like(p) for p in fine.ordered_parameters()]
            garr = tn._lib.pointer_array(grads)
            tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), SC + SF, m,
                                                    P(sig), P(rgb), P(cache, torch.uint8), tn._lib.stream()), "fwd-train")

            tiles = (m + 127) // 128

            def k_bwd(phases=7):
                tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_s),
                                                              P(g_c), garr, P(scratch, torch.uint8), phases, 0, tiles, 0,
                                                              tn._lib.stream()), "bwd")
            k_bwd()
            sec_w = timed_kernel(lambda: k_bwd(4))  # the weight-gradient kernel alone, on the scratch a full backward filled
            alg_bytes = tiles * (83 * 16384 + 2 * 128 * 16)  # 83 blocks of 16 KB (G and X tile images) + head gradients per tile
            ach_w = alg_bytes / sec_w / 1e9
            roof = {"bound": "hbm", "kernel": f"mlp_wgrad_kernel (split-K tcgen05 weight gradients, {m} rows)",
                    "achieved": ach_w, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_w / pk["hbm_gbs"],
                    "traffic": recorded_traffic("mlp_wgrad_kernel", m),  # dram bytes of one launch from ncu --set full, or null
                    "algorithmic_bytes": alg_bytes, "peak_source": pk["src"] + " (burst: kernel timed alone)",
                    "peak_note": "the peak is a copy (read+write) figure; this kernel only reads, and a read-only stream "
                                 "measures 6.7 TB/s on the same pool (tools/prof_store.py), hence frac slightly above 1",
                    "launch_ms": sec_w * 1e3}
            del cache, scratch
    # ---- full 800x800 frame (configs[2]): every rank renders its contiguous share of the pixels (no collective in the
    #      data path); the frame time is the slowest rank's, measured with CUDA events between barriers
    if args.precision == "bf16" or rank == 0:
        from torch_nerf_b200.parallel import shard_range

        cam = cams[0]
        lo, hi = shard_range(IMG * IMG, rank, world) if args.precision == "bf16" else (0, IMG * IMG)
        for _ in range(2):
            eng_render.render_frame(cam, False, lo, hi - lo)
        torch.cuda.synchronize()
        if world > 1 and args.precision == "bf16":
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ev0.record()
        for _ in range(reps):
            img = eng_render.render_frame(cam, False, lo, hi - lo)
        ev1.record()
        torch.cuda.synchronize()
        frame_ms = ev0.elapsed_time(ev1) / reps
        if world > 1 and args.precision == "bf16":
            tmax = torch.tensor([frame_ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            frame_ms = float(tmax.item())
        flop_frame = IMG * IMG * (SC + SC + SF) * FLOP_FWD_PER_EVAL
        sharded = world if args.precision == "bf16" else 1
        render = {"frame_ms": frame_ms, "n_gpus": sharded, "resolution": [IMG, IMG], "rays_per_s": IMG * IMG / (frame_ms * 1e-3),
                  "tensor_frac_sustained": flop_frame / (frame_ms * 1e-3) / 1e12 / (pk["tf_sustained"] * sharded),
                  "finite": bool(torch.isfinite(img).all().item())}
        if sharded == 1:
            render["frame_ms_1gpu"] = frame_ms
        # configs[4]: LLFF-shaped forward-facing view, 1008x756, NDC rays, near/far 0/1 (runner_utils.py:489-491), same
        # sharding; two frames timed
        if args.precision == "bf16":
            lw, lh, lf = 1008, 756, 815.0
            c2w = torch.eye(4)[:3, :4].clone()
            c2w[:, 3] = torch.tensor([0.05, -0.02, 0.1])
            cam_l = tn.PerspectiveCamera({"f_x": lf, "f_y": lf, "img_width": lw, "img_height": lh}, c2w, 0.0, 1.0)
            lo_l, hi_l = shard_range(lw * lh, rank, world)
            eng_render.render_frame(cam_l, True, lo_l, hi_l - lo_l)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0.record()
            for _ in range(2):
                img_l = eng_render.render_frame(cam_l, True, lo_l, hi_l - lo_l)
            ev1.record()
            torch.cuda.synchronize()
            llff_ms = ev0.elapsed_time(ev1) / 2
            if world > 1:
                tmax = torch.tensor([llff_ms], device=dev)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                llff_ms = float(tmax.item())
            render["llff_1008x756_ndc"] = {"frame_ms": llff_ms, "rays_per_s": lw * lh / (llff_ms * 1e-3),
                                           "finite": bool(torch.isfinite(img_l).all().item())}

    # ---- HBM-bound stages: ray generation, sampling, compositing at the frame's ray count, each kernel timed
    #      alone with CUDA events; algorithmic bytes per unit as listed in DESIGN.md section 4
    stages = None
    if rank == 0 and args.precision == "bf16":
        time.sleep(3.0)  # as above: each stage kernel is timed alone
        lib = tn._lib.load()
        P, st = tn._lib.ptr, tn._lib.stream
        nr = IMG * IMG
        S = SC + SF
        cam_s = cams[0].pack(False)
        ro = torch.empty(nr, 3, device=dev); rd = torch.empty(nr, 3, device=dev)
        u_c = torch.rand(nr, SC, device=dev); u1 = torch.rand(nr, SF, device=dev); u2 = torch.rand(nr, SF, device=dev)
        w_c = torch.rand(nr, SC, device=dev)
        t_c = torch.empty(nr, SC, device=dev); d_c = torch.empty(nr, SC, device=dev)
        t_f = torch.empty(nr, S, device=dev); d_f = torch.empty(nr, S, device=dev)
        sig = torch.rand(nr, S, device=dev); rad = torch.rand(nr, S, 3, device=dev)
        rgb_o = torch.empty(nr, 3, device=dev); w_o = torch.empty(nr, S, device=dev)
        g_rgb = torch.rand(nr, 3, device=dev); g_sig = torch.empty(nr, S, device=dev); g_rad = torch.empty(nr, S, 3, device=dev)

        def timed(fn, reps=5):
            for _ in range(2):
                fn()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) * 1e-3 / reps

        ck = tn._lib.check
        runs = [
            ("raygen_kernel", nr * 24,
             lambda: ck(lib.nerf_generate_rays_from_pixels(None, 0, nr, cam_s, P(ro), P(rd), st()), "raygen")),
            ("sample_coarse_flat_kernel", nr * SC * 12,
             lambda: ck(lib.nerf_sample_coarse(P(ro), P(rd), nr, SC, 2.0, 6.0, P(u_c), P(t_c), None, None, P(d_c), st()), "coarse")),
            ("sample_fine_64_128_kernel", nr * (SC * 8 + SF * 8 + S * 8),
             lambda: ck(lib.nerf_sample_fine(P(ro), P(rd), nr, SC, SF, 2.0, 6.0, P(w_c), P(u_c), P(u1), P(u2), None, P(t_f), None,
                                             None, P(d_f), st()), "fine")),
            ("composite_fwd_reg_kernel<6>", nr * S * 24 + nr * 12,
             lambda: ck(lib.nerf_composite_fwd(P(sig), P(rad), P(d_f), None, nr, S, P(rgb_o), P(w_o), None, None, st()), "comp")),
            ("composite_bwd_blk_kernel<6>", nr * S * 36 + nr * 12,
             lambda: ck(lib.nerf_composite_bwd(P(sig), P(rad), P(d_f), P(g_rgb), None, nr, S, P(g_sig), P(g_rad), st()), "compb")),
        ]
        stages = []
        for name, nbytes, fn in runs:
            sec = timed(fn)
            stages.append({"kernel": name, "rays": nr, "bound": "hbm", "algorithmic_bytes": nbytes, "launch_ms": sec * 1e3,
                           "achieved": nbytes / sec / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": nbytes / sec / 1e9 / pk["hbm_gbs"]})
        del ro, rd, u_c, u1, u2, w_c, t_c, d_c, t_f, d_f, sig, rad, rgb_o, w_o, g_rgb, g_sig, g_rad

    if rank == 0:
        cpu = None
        ref_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_block(args.quick)
        if world == 1 and not args.no_reference_gpu:
            try:
                ref_gpu = reference_gpu_block(dev, n_rays)
                ref_gpu["speedup_train"] = value / ref_gpu["train"]["rays_per_s"]
                if render is not None:
                    ref_gpu["speedup_render_800x800"] = ref_gpu["render_800x800_ms"] / render["frame_ms"]
            except Exception as err:  # the comparator must never take the bench line down with it
                ref_gpu = {"unavailable": f"{type(err).__name__}: {err}"[:300]}
        dropin = None
        if world == 1 and args.precision == "bf16" and not args.no_dropin:
            try:
                dropin = dropin_block(tn, dev, n_rays, eng_render)
            except Exception as err:
                dropin = {"unavailable": f"{type(err).__name__}: {err}"[:300]}
        if isinstance(ref_gpu, dict) and isinstance(dropin, dict) and "c1_render_100x100_ms" in ref_gpu:
            if "c1_render_100x100_engine_bf16_ms" in dropin:  # config C1 (100x100 frame): reference on torch-CUDA / this path
                ref_gpu["speedup_c1_render_100x100"] = ref_gpu["c1_render_100x100_ms"] / dropin["c1_render_100x100_engine_bf16_ms"]
            if "train_bf16" in dropin:  # and the drop-in API alone (classes swapped, reference's own loop otherwise)
                ref_gpu["speedup_train_dropin_api"] = ref_gpu["train"]["ms_per_step"] / dropin["train_bf16"]["ms_per_step"]
        step_flop = n_rays * (SC + SC + SF) * FLOP_TRAIN_PER_EVAL
        # the whole step against both roofs: its tensor work against the sustained bf16 peak, and the HBM bytes its three
        # tensor-core kernel families move (ncu dram bytes of the fine pass's launches, recorded in profiles/traffic.json,
        # x 4/3 for the coarse pass) against the copy bandwidth -- the design stages activations and activation
        # gradients in HBM, so the step sits between the two roofs
        roof_step = None
        if args.precision == "bf16":
            m_f = n_rays * (SC + SF)
            tr = [recorded_traffic(k, m_f) for k in ("mlp_fwd_kernel<1>", "mlp_dgrad_kernel", "mlp_wgrad_kernel")]
            hbm_bytes = None if any(t is None for t in tr) else sum(tr) * (SC + SC + SF) / (SC + SF)
            sec_step = ms_dev / args.steps * 1e-3
            roof_step = {"tensor": {"achieved": step_flop / sec_step / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                                    "frac": step_flop / sec_step / 1e12 / pk["tf_sustained"]},
                         "hbm": None if hbm_bytes is None else
                         {"achieved": hbm_bytes / sec_step / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": hbm_bytes / sec_step / 1e9 / pk["hbm_gbs"], "traffic": hbm_bytes},
                         "peak_source": pk["src"] + " (sustained: kernels timed inside the step)"}
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_rays else "weak",
            "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"single training step (coarse 64 + fine 64+128, fwd+bwd, Adam), {n_rays}-ray batch per GPU, "
                                   "lego-shaped synthetic scene, 800x800 cameras, near/far 2/6",
                       "rays_per_gpu": n_rays, "global_rays": n_rays * world, "samples": [SC, SF],
                       "l2_policy": "per-step working set (activation cache + gradients, >1 GB) exceeds the 126 MB L2",
                       "precision": args.precision, "cuda_graph": bool(use_graph),
                       "exchange": "none (1 GPU)" if world == 1 else
                                   ("peer-memory reduce + Adam in one kernel (csrc/dp_exchange.cu)" if px is not None
                                    else "NCCL all-reduce + Adam")},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 8 + n_rays * 12,
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks,
            "step_tensor_frac_sustained": step_flop / (ms_dev / args.steps * 1e-3) / 1e12 / pk["tf_sustained"],
            "roofline": roof, "roofline_mlp": roof_mlp, "roofline_step": roof_step, "roofline_stages": stages, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "dropin": dropin, "dp_parity": parity,
            "c4_global_32768": c4,
            "render": render,
            "loss_last": [float(x) for x in losses_host[total_steps - 1]],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NERF_B200_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per GPU (weak scaling)")
    ap.add_argument("--global-rays", type=int, default=0, help="global batch split over the GPUs (config C4: 32768); strong scaling")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra 32768-ray global-batch timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the same-box torch-CUDA run of the unmodified reference")
    ap.add_argument("--no-dropin", action="store_true", help="skip timing the reference-facing plugin API (VolumeRenderer.render_scene)")
    ap.add_argument("--quick", action="store_true", help="shorter reference legs")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel of the iteration instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
