"""Two threads issuing tcgen05.mma into ONE accumulator (K steps split between them): is the result still exact?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
torch.manual_seed(0)
worst = 0.0
bad = 0
for trial in range(300):
    n, k = (128, 256) if trial % 2 == 0 else (256, 256)
    a = torch.randn(128, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    d = torch.zeros(128, n, device="cuda")
    tn._lib.check(tn._lib.load_selftest().nerf_selftest_umma(VP(a.data_ptr()), VP(b.data_ptr()), P(d), n, k, 3, tn._lib.stream()), "umma split")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T
    err = (d - ref).abs().max().item()
    worst = max(worst, err)
    bad += err > 1e-3
print(f"300 trials, two issuing threads into one accumulator: worst |err| {worst:.3e}, trials above 1e-3: {bad}")

# exact-arithmetic stress: operands in {-1, 0, 1}, the product accumulated 64 times by the two threads -> every partial
# sum is an integer below 2^24, so ONE lost or doubled MMA shows as an exact mismatch
mism = 0
for trial in range(100):
    n, k, reps = 256, 256, 64
    a = torch.randint(-1, 2, (128, k), device="cuda").bfloat16(); b = torch.randint(-1, 2, (n, k), device="cuda").bfloat16()
    d = torch.zeros(128, n, device="cuda")
    tn._lib.check(tn._lib.load_selftest().nerf_selftest_umma(VP(a.data_ptr()), VP(b.data_ptr()), P(d), n, k, 3 + 16 * (reps - 1), tn._lib.stream()), "umma split")
    torch.cuda.synchronize()
    ref = (a.float() @ b.float().T) * reps
    mism += int((d != ref).sum().item())
print(f"100 trials x 64 repetitions x 16 MMAs, integer operands: {mism} mismatching accumulator entries")
