"""CTA-pair MMA (tcgen05 cta_group::2): correctness against a plain matmul, then the sustained issue rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
torch.manual_seed(0)
for ts in (0, 1):
    for n in (64, 128, 256):
        for k in (64, 256):
            a = torch.randn(256, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
            d = torch.zeros(256, n, device="cuda")
            tn._lib.check(tn._lib.load_selftest().nerf_selftest_umma2(VP(a.data_ptr()), VP(b.data_ptr()), P(d), n, k, ts, 1, 0, None, tn._lib.stream()), "umma2")
            torch.cuda.synchronize()
            ref = a.float() @ b.float().T
            err = (d - ref).abs().max().item()
            print(f"ts={ts} n={n:3d} k={k:3d}: max|err| = {err:.3e}  (ref max {ref.abs().max().item():.1f})")
cyc = torch.zeros(74, dtype=torch.int64, device="cuda")
for ts in (0, 1):
    for n in (128, 256):
        k = 256
        a = torch.randn(256, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
        d = torch.zeros(256, n, device="cuda")
        iters = 500
        tn._lib.check(tn._lib.load_selftest().nerf_selftest_umma2(VP(a.data_ptr()), VP(b.data_ptr()), P(d), n, k, ts, 74, iters, VP(cyc.data_ptr()), tn._lib.stream()), "umma2")
        torch.cuda.synchronize()
        per = cyc.float().mean().item() / (iters * k / 16)
        print(f"rate {'TS' if ts else 'SS'} pair M=256 N={n}: {per:6.1f} cycles per MMA (tensor floor {n / 2:.0f}; single-CTA M=128 measured: TS 93/137, SS 104/168)")

# both CTAs of a pair issuing pair MMAs concurrently, each into its own accumulator: each instruction feeds BOTH SMs' tensor
# pipes, so two issuers offer every pipe 2 x (n/2) cycles of work per instruction time
cyc2 = torch.zeros(148, dtype=torch.int64, device="cuda")
for ts in (0, 1):
    for n in (64, 128):
        k = 256
        a = torch.randn(256, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
        d = torch.zeros(256, n, device="cuda")
        iters = 500
        tn._lib.check(tn._lib.load_selftest().nerf_selftest_umma2(VP(a.data_ptr()), VP(b.data_ptr()), P(d), n, k, ts | 2, 74, iters, VP(cyc2.data_ptr()), tn._lib.stream()), "umma2 both")
        torch.cuda.synchronize()
        err = (d - a.float() @ b.float().T).abs().max().item()
        per = cyc2.float().mean().item() / (2 * iters * k / 16)
        print(f"rate {'TS' if ts else 'SS'} pair M=256 N={n}, BOTH CTAs issuing: {per:6.1f} cycles per MMA aggregate (tensor floor {n / 2:.0f}); "
              f"second CTA's accumulator max|err| = {err:.3e}")
