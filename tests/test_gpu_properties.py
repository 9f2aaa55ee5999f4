"""Size-independent properties at BASELINE.json's full sizes (no oracle needed): the domain's own invariants.

  * the tensor-core MLP is row-wise: permuting the sample rows of a 786 432-row batch (the fine pass of a 4096-ray step)
    permutes the outputs BIT-EXACTLY -- whatever tile, slot and CTA a row lands in (cube.py:39-76 flattens rows; nerf.py is
    row-wise);
  * compositing is linear in the radiance and its weights partition unity: rgb(a c1 + b c2) = a rgb(c1) + b rgb(c2),
    sum_i w_i = 1 - exp(-sum_i sigma_i delta_i) (quadrature_integrator.py:41-65), on a full 800x800 frame of rays;
  * the backward is linear in the upstream gradient: grads(g1 + g2) = grads(g1) + grads(g2) up to fp32 accumulation order.
"""
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


def test_mlp_chain_is_row_permutation_equivariant_at_full_size(tn):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(1)
    m = 4096 * 192
    net = tn.NeRF(63, 27, precision="bf16").cuda()
    packed = net.packed_weights(True)
    pts = (torch.rand(m, 3, device="cuda") - 0.5) * 8
    dirs = torch.randn(m, 3, device="cuda")
    perm = torch.randperm(m, device="cuda")

    def run(p, d):
        sig, rgb = torch.empty(m, device="cuda"), torch.empty(m, 3, device="cuda")
        tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), P(p), P(d), None, None, None, 0, m, P(sig), P(rgb), None, st()),
                      "fwd")
        return sig, rgb

    s0, r0 = run(pts, dirs)
    s1, r1 = run(pts[perm].contiguous(), dirs[perm].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(s0[perm], s1) and torch.equal(r0[perm], r1)
    assert bool(torch.isfinite(s0).all()) and float(r0.min()) >= 0.0 and float(r0.max()) <= 1.0


def test_compositing_linearity_and_partition_of_unity_full_frame(tn):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(2)
    n, s = 800 * 800, 192
    sigma = torch.rand(n, s, device="cuda") * 2
    delta = torch.rand(n, s, device="cuda") * 0.04
    c1, c2 = torch.rand(n, s, 3, device="cuda"), torch.rand(n, s, 3, device="cuda")
    a, b = 0.25, 1.5

    def comp(c):
        rgb, w = torch.empty(n, 3, device="cuda"), torch.empty(n, s, device="cuda")
        tn._lib.check(lib.nerf_composite_fwd(P(sigma), P(c), P(delta), None, n, s, P(rgb), P(w), None, None, st()), "comp")
        return rgb, w

    r1, w1 = comp(c1)
    r2, w2 = comp(c2)
    r12, _ = comp((a * c1 + b * c2).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(w1, w2)                                         # the weights do not depend on the radiance
    torch.testing.assert_close(r12, a * r1 + b * r2, rtol=1e-5, atol=2e-6)
    total = torch.exp(-(sigma.double() * delta.double()).sum(-1))
    torch.testing.assert_close(w1.double().sum(-1), 1.0 - total, rtol=0, atol=2e-5)
    assert float(w1.min()) >= 0.0


def test_backward_is_linear_in_the_upstream_gradient(tn):
    lib, P, st = tn._lib.load(), tn._lib.ptr, tn._lib.stream
    torch.manual_seed(3)
    n, s = 512, 192
    m = n * s
    net = tn.NeRF(63, 27, precision="bf16").cuda()
    packed = net.packed_weights(True)
    ray_o, ray_d = torch.randn(n, 3, device="cuda"), torch.randn(n, 3, device="cuda")
    t = torch.rand(n, s, device="cuda") * 4 + 2
    sig, rgb = torch.empty(m, device="cuda"), torch.empty(m, 3, device="cuda")
    cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(lib.nerf_mlp_bf16_bwd_scratch_bytes(m), dtype=torch.uint8, device="cuda")
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(cache, torch.uint8), st()), "fwd")

    def grads(gs, gc):
        out = [torch.empty_like(p) for p in net.ordered_parameters()]
        tn._lib.check(lib.nerf_mlp_bf16_backward(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(gs), P(gc),
                                                 tn._lib.pointer_array(out), P(scratch, torch.uint8), st()), "bwd")
        torch.cuda.synchronize()
        return torch.cat([g.reshape(-1) for g in out]).double()

    # powers of two keep the bf16 rounding of the activation gradients identical between the runs: G(2 g) = 2 G(g) exactly
    gs, gc = torch.randn(m, device="cuda") * 1e-3, torch.randn(m, 3, device="cuda") * 1e-3
    g1 = grads(gs, gc)
    g2 = grads((2 * gs).contiguous(), (2 * gc).contiguous())
    scale = float(g1.abs().max())
    assert float((g2 - 2 * g1).abs().max()) <= 2e-5 * scale   # only the order of the fp32 atomics differs
    gz = grads(torch.zeros_like(gs), torch.zeros_like(gc))
    assert float(gz.abs().max()) == 0.0
