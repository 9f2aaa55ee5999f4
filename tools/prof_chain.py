"""MMA issue-rate micro-benchmark + in-kernel timeline of the forward chain (prints a small report)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn

lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
def rate(blocks, iters, n, mode, bg_warps=0, bg_iters=0, bg_store=0):
    out.zero_()
    tn._lib.check(tn._lib.load_selftest().nerf_selftest_mma_rate(blocks, iters, n, mode, bg_warps, bg_iters, bg_store, VP(out.data_ptr()), tn._lib.stream()), "rate")
    torch.cuda.synchronize()
    mma = out[:blocks].float().mean().item() / max(1, iters * 4)
    bg = out[blocks:2 * blocks].float().mean().item() / max(1, bg_iters)
    return mma, bg
if "--micro" in sys.argv:
    for mode, n in ((0, 256), (0, 128), (1, 256), (1, 128)):
        m, _ = rate(148, 2000, n, mode)
        print(f"mma alone   mode={'SS' if mode == 0 else 'TS'} N={n:3d}: {m:6.1f} cycles/MMA (floor {128 * n / 256:.0f})")
    for n in (128, 256):
        for flags, label in ((2, "commit per 4"), (4, "probe per 4"), (6, "commit+probe per 4"), (8, "syncwarp per 4"), (14, "all three")):
            m, _ = rate(148, 2000, n, 1 | flags)
            print(f"mma TS N={n:3d} + {label:20s}: {m:6.1f} cycles/MMA")
    for mode, n in ((16, 128), (17, 128), (17, 64)):
        m, _ = rate(148, 2000, n, mode)
        print(f"mma alternating two accumulators mode={'SS' if (mode & 1) == 0 else 'TS'} N={n:3d}: {m:6.1f} cycles/MMA (floor {128 * n / 256:.0f})")
    for mode, n in ((1 | 4, 128), (1 | 4 | 64, 128)):
        m, _ = rate(148, 2000, n, mode)
        print(f"mma TS N={n} + probe per 4 with {'test_wait' if mode & 64 else 'try_wait'}: {m:6.1f} cycles/MMA")
    for mode, n in ((33 | 128, 128), (33 | 128, 256)):
        m, _ = rate(148, 2000, n, mode)
        print(f"two issuer warps, ONE accumulator, TS N={n:3d}: {m / 2:6.1f} cycles/MMA aggregate")
    for mode, n in ((33, 128), (33, 64), (32, 128)):
        m, _ = rate(148, 2000, n, mode)
        print(f"two issuer warps, own accumulators, mode={'SS' if (mode & 1) == 0 else 'TS'} N={n:3d}: {m / 2:6.1f} cycles/MMA aggregate ({m:6.1f} per issuer)")
    for mode, n in ((256, 128), (257, 128), (257 | 32, 128), (257, 256)):
        m, _ = rate(148, 2000, n, mode)
        two = bool(mode & 32)
        print(f"mma M=64 mode={'TS' if mode & 1 else 'SS'} N={n:3d}{' two issuers' if two else ''}: {m / (2 if two else 1):6.1f} cycles/MMA"
              f"{' aggregate' if two else ''} (an M=128 instruction of this N: {128 * n / 256:.0f} pipe cycles)")
    if "--short" in sys.argv:
        sys.exit(0)
    for w in (1, 4, 8, 16):
        _, l = rate(148, 0, 128, 0, w, 4000, 0)
        _, st = rate(148, 0, 128, 0, w, 4000, 1)
        print(f"tmem alone  {w:2d} warps: ld {l:6.1f} cycles per 32x32 fp32 load per warp -> {w * 4096 / l:6.1f} B/cycle/SM ; "
              f"st {st:6.1f} cycles -> {w * 4096 / st:6.1f} B/cycle/SM")
    for mode, n in ((0, 256), (1, 128), (0, 128)):
        for w in (4, 8):
            for store in (0, 1):
                m, l = rate(148, 2000, n, mode, w, 100000, store)  # background runs longer than the MMAs
                m2, l2 = rate(148, 2000, n, mode, w, 2000, store)
                print(f"mma + {w} warps of tmem {'st' if store else 'ld'}  mode={'SS' if mode == 0 else 'TS'} N={n:3d}: {m:6.1f} cycles/MMA")
    sys.exit(0)

n, s = 4096, 192
m = n * s
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights()
ray_o = torch.randn(n, 3, device="cuda"); ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda"); rgb = torch.empty(m, 3, device="cuda")
tiles = 4
prof = torch.zeros(tiles * 10 * 8, dtype=torch.int64, device="cuda")
def fwd(cache=None):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(cache, torch.uint8) if cache is not None else None, tn._lib.stream()), "fwd")
train_cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
cases = [("inference", None, 0), ("training", train_cache, 0)]
if "--nostore" in sys.argv:
    cases.append(("training, block stores skipped (debug)", train_cache, 1))
for label, cache, mode in cases:
    fwd(cache); torch.cuda.synchronize()
    lib.nerf_debug_set_profile_buffer(VP(prof.data_ptr()), tiles | (mode << 16))
    prof.zero_()
    fwd(cache); torch.cuda.synchronize()
    lib.nerf_debug_set_profile_buffer(None, mode << 16)
    p = prof.cpu().view(tiles, 10, 8)
    t0 = int(p[0, 0, 0])
    print(f"--- forward chain timeline ({label}), CTA 0, cycles relative to first layer start")
    print("tile layer  mma_start  mma_issued  acc_seen  epi_done | mma_span  epi_span  layer_period  wait_act  wait_w  st_wait  st_rest")
    prev = None
    for ti in range(1, tiles):
        for l in range(10):
            a, b, c, d = [int(x) - t0 for x in p[ti, l, :4]]
            period = (a - prev) if prev is not None else 0
            prev = a
            print(f"{ti:4d} {l:5d} {a:10d} {b:11d} {c:9d} {d:9d} | {b - a:8d} {d - c:9d} {period:9d} {int(p[ti, l, 4]):9d} {int(p[ti, l, 5]):7d} {int(p[ti, l, 6]):8d} {int(p[ti, l, 7]):8d}")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): fwd(cache)
    ev1.record(); torch.cuda.synchronize()
    print(f"{label}: {ev0.elapsed_time(ev1)/5*1e3:.0f} us per launch of {m} rows")
    lib.nerf_debug_set_profile_buffer(None, 0)
