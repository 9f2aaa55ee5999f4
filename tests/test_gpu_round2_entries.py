"""GPU tests of the C-ABI entry points added in round 2, each against the entry points it replaces / composes:

    nerf_mlp_bf16_backward_part   pieces (phase mask, tile ranges, CTA counts) add up to nerf_mlp_bf16_backward
    nerf_composite_bwd_mse        == nerf_mse_loss + nerf_composite_bwd                (train.py:180/202, runner_utils.py:731)
    nerf_train_prologue           == 2 x nerf_mlp_bf16_pack + zero fills               (train.py:131, runner_utils.py:569-660)
    nerf_dp_exchange_adam         world = 1: == nerf_adam_step; exchange-only mode leaves the buffer unchanged
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import torch_nerf_b200 as mod

    mod._lib.load()
    return mod


def test_backward_in_pieces_adds_up(tn):
    lib, P = tn._lib.load(), tn._lib.ptr
    torch.manual_seed(0)
    n, s = 300, 192                      # 57600 rows = 450 tiles
    m = n * s
    tiles = (m + 127) // 128
    net = tn.NeRF(63, 27, precision="bf16").cuda()
    packed = net.packed_weights(True)
    ray_o, ray_d = torch.randn(n, 3, device="cuda"), torch.randn(n, 3, device="cuda")
    t = torch.rand(n, s, device="cuda") * 4 + 2
    sig, rgb = torch.empty(m, device="cuda"), torch.empty(m, 3, device="cuda")
    cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(lib.nerf_mlp_bf16_bwd_scratch_bytes(m), dtype=torch.uint8, device="cuda")
    st = tn._lib.stream
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(cache, torch.uint8), st()), "fwd")
    g_s, g_c = torch.randn(m, device="cuda") * 1e-3, torch.randn(m, 3, device="cuda") * 1e-3

    def grads_of(run):
        grads = [torch.full_like(p, 7.0) for p in net.ordered_parameters()]  # garbage: phase 1 must zero it
        run(tn._lib.pointer_array(grads))
        torch.cuda.synchronize()
        return grads

    whole = grads_of(lambda ga: tn._lib.check(lib.nerf_mlp_bf16_backward(
        P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_s), P(g_c), ga, P(scratch, torch.uint8), st()), "bwd"))

    def pieces(ga):
        part = lambda ph, a, b, ctas: tn._lib.check(lib.nerf_mlp_bf16_backward_part(
            P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_s), P(g_c), ga, P(scratch, torch.uint8), ph, a, b, ctas,
            st()), "part")
        part(1, 0, 0, 0)                       # zero only (empty range)
        bounds = [0, 100, 101 * 2, tiles]      # even starts, uneven sizes
        for (a, b), ctas in zip(zip(bounds[:-1], bounds[1:]), (0, 37, 90)):
            part(2, a, b, ctas)                # chain of this range ...
            part(4, a, b, 148 - ctas if ctas else 0)  # ... then its weight gradients, accumulated
    split = grads_of(pieces)
    for a, b in zip(split, whole):
        scale = float(b.abs().max()) + 1e-20
        assert float((a - b).abs().max()) / scale < 2e-4
    with pytest.raises(ValueError):   # odd tile0
        tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_s), P(g_c),
                                                      tn._lib.pointer_array(whole), P(scratch, torch.uint8), 7, 1, tiles, 0, st()), "part")


@pytest.mark.parametrize("s", [64, 192, 100])
def test_composite_bwd_with_folded_mse(tn, s):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(s)
    n = 1000
    sigma, rad = torch.rand(n, s, device="cuda") * 3, torch.rand(n, s, 3, device="cuda")
    delta = torch.rand(n, s, device="cuda") * 0.05
    delta[:, -1] = 1e8
    target = torch.rand(n, 3, device="cuda")
    rgb, w = torch.empty(n, 3, device="cuda"), torch.empty(n, s, device="cuda")
    tn._lib.check(lib.nerf_composite_fwd(P(sigma), P(rad), P(delta), None, n, s, P(rgb), P(w), None, None, st()), "fwd")
    g_rgb, loss_a = torch.empty(n, 3, device="cuda"), torch.zeros(1, device="cuda")
    gs_a, gr_a = torch.empty_like(sigma), torch.empty_like(rad)
    tn._lib.check(lib.nerf_mse_loss(P(rgb), P(target), n, P(g_rgb), P(loss_a), st()), "mse")
    tn._lib.check(lib.nerf_composite_bwd(P(sigma), P(rad), P(delta), P(g_rgb), None, n, s, P(gs_a), P(gr_a), st()), "bwd")
    loss_b = torch.zeros(1, device="cuda")
    gs_b, gr_b = torch.empty_like(sigma), torch.empty_like(rad)
    tn._lib.check(lib.nerf_composite_bwd_mse(P(sigma), P(rad), P(delta), P(rgb), P(target), n, s, P(gs_b), P(gr_b), P(loss_b), st()),
                  "bwd_mse")
    torch.cuda.synchronize()
    assert torch.equal(gs_a, gs_b) and torch.equal(gr_a, gr_b)          # same arithmetic on the same g_rgb values
    assert abs(float(loss_a) - float(loss_b)) <= 1e-6 * abs(float(loss_a))
    assert abs(float(loss_a) - float(torch.mean((rgb - target) ** 2))) <= 1e-5 * float(loss_a)


def test_train_prologue_equals_pack_and_zero(tn):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(3)
    a, b = tn.NeRF(63, 27, precision="bf16").cuda(), tn.NeRF(63, 27, precision="bf16").cuda()
    ref_a, ref_b = a.packed_weights(True).clone(), b.packed_weights(True).clone()
    nb = lib.nerf_mlp_bf16_packed_bytes()
    pa = torch.full((nb,), 0x5A, dtype=torch.uint8, device="cuda")
    pb = torch.full((nb,), 0x5A, dtype=torch.uint8, device="cuda")
    z0, z1 = torch.full((1191688,), 3.0, device="cuda"), torch.full((2,), 5.0, device="cuda")
    arr = lambda net: tn._lib.pointer_array([p.detach() for p in net.ordered_parameters()])
    tn._lib.check(lib.nerf_train_prologue(arr(a), P(pa, torch.uint8), arr(b), P(pb, torch.uint8), P(z0), z0.numel(), P(z1), 2, st()),
                  "prologue")
    torch.cuda.synchronize()
    assert torch.equal(pa, ref_a) and torch.equal(pb, ref_b)
    assert float(z0.abs().max()) == 0.0 and float(z1.abs().max()) == 0.0
    # empty zero regions are allowed
    tn._lib.check(lib.nerf_train_prologue(arr(a), P(pa, torch.uint8), arr(b), P(pb, torch.uint8), None, 0, None, 0, st()), "prologue")


def test_dp_exchange_single_rank_equals_adam(tn):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(9)
    n = 2 * 595844
    p0 = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 1e-2
    runs = []
    for fused in (False, True):
        p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        grad = g.clone()
        flags, counter = torch.zeros(64, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
        gp = (ctypes.c_void_p * 1)(grad.data_ptr())
        fp = (ctypes.c_void_p * 1)(flags.data_ptr())
        for step in (1, 2, 3):
            if fused:
                tn._lib.check(lib.nerf_dp_exchange_adam(gp, fp, 0, 1, P(p), P(m), P(v), n, 5e-4, 0.9, 0.999, 1e-8, step, 0.5, step,
                                                        P(counter, torch.int32), st()), "dp")
            else:
                tn._lib.check(lib.nerf_adam_step(P(p), P(grad), P(m), P(v), n, 5e-4, 0.9, 0.999, 1e-8, step, 0.5, st()), "adam")
        torch.cuda.synchronize()
        runs.append((p, m, v, grad))
    for x, y in zip(runs[0][:3], runs[1][:3]):
        assert torch.equal(x, y)
    assert torch.equal(runs[1][3], g)  # world = 1: the "sum over ranks" is the buffer itself
    # exchange-only mode (param NULL) must not touch anything but the (unchanged) gradient buffer
    flags, counter = torch.zeros(64, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    grad = g.clone()
    tn._lib.check(lib.nerf_dp_exchange_adam((ctypes.c_void_p * 1)(grad.data_ptr()), (ctypes.c_void_p * 1)(flags.data_ptr()), 0, 1,
                                            None, None, None, n, 0.0, 0.9, 0.999, 1e-8, 1, 1.0, 1, P(counter, torch.int32), st()), "dp")
    torch.cuda.synchronize()
    assert torch.equal(grad, g) and int(flags[0]) == 1 and int(flags[1]) == 1 and int(counter) == 0
