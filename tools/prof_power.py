"""Which kernel runs into the board's power cap?  Each hot kernel is launched back to back for ~0.7 s while NVML is
sampled every 5 ms (SM clock, board power); printed: time per launch in that sustained loop, median / minimum SM clock,
median power.  (The training step as a whole shows `sw_power_cap` and SM clocks down to ~1.6 GHz.)"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pynvml
import torch
import torch_nerf_b200 as tn

lib = tn._lib.load()
P = tn._lib.ptr
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n, s = 4096, 192
m = n * s
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights()
ray_o = torch.randn(n, 3, device="cuda"); ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda"); rgb = torch.empty(m, 3, device="cuda")
cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
scratch = torch.empty(lib.nerf_mlp_bf16_bwd_scratch_bytes(m), dtype=torch.uint8, device="cuda")
g_sig = torch.randn(m, device="cuda") * 1e-3; g_rgb = torch.randn(m, 3, device="cuda") * 1e-3
grads = [torch.zeros_like(p) for p in net.ordered_parameters()]
garr = tn._lib.pointer_array(grads)

def fwd(c):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(c, torch.uint8) if c is not None else None, tn._lib.stream()), "fwd")
def bwd(mask):
    tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_sig), P(g_rgb), garr,
                                                  P(scratch, torch.uint8), mask, 0, (m + 127) // 128, 0, tn._lib.stream()), "bwd")
cases = [("mlp_fwd inference", lambda: fwd(None)), ("mlp_fwd training", lambda: fwd(cache)), ("mlp_dgrad", lambda: bwd(2)),
         ("mlp_wgrad", lambda: bwd(4))]
fwd(cache); bwd(7); torch.cuda.synchronize()
for name, fn in cases:
    time.sleep(2.0)
    samples, stop = [], threading.Event()
    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            stop.wait(0.005)
    reps = 600
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th = threading.Thread(target=sampler); th.start()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    stop.set(); th.join()
    us = e0.elapsed_time(e1) * 1e3 / reps
    clk = np.array([c for c, _ in samples[len(samples) // 4:]]); pw = np.array([p for _, p in samples[len(samples) // 4:]])
    print(f"{name:20s} {us:8.1f} us/launch sustained | SM clock median {np.median(clk):6.0f} min {clk.min():6.0f} MHz | power median {np.median(pw):6.0f} W max {pw.max():6.0f} W")
