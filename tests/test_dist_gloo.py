"""CPU, world_size 2 over gloo: the ray-sharded data-parallel step (torch-nerf_b200/parallel.py).  Each rank computes
the gradients of its ray shard (with the numpy oracle standing in for the GPU kernels), the flat gradient buffers
are all-reduced and averaged, and the result must equal the single-process gradient of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    from oracle import nerf_oracle as orc

    rng = np.random.default_rng(0)
    n = 8
    pc, pf = orc.init_nerf_params(seed=1), orc.init_nerf_params(seed=2)
    focal = orc.blender_focal(800)
    c2w = orc.pose_spherical(30.0, -30.0, 4.0)
    pix = rng.choice(800 * 800, size=n, replace=False).astype(np.int64)
    target = rng.random((n, 3), dtype=np.float32)
    u = [rng.random((n, k), dtype=np.float32) for k in (16, 16, 32, 32)]
    return orc, pc, pf, focal, c2w, pix, target, u


def _flat_grads(orc, pc, pf, focal, c2w, pix, target, u):
    coords = orc.screen_coords(800, 800)[pix]
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, 800, 800), c2w, 2.0, 800, 800, False)
    out = orc.train_step_grads(pc, pf, o, d, 2.0, 6.0, 16, 32, target, *u)
    keys = list(pc.keys())
    return np.concatenate([out["coarse_grads"][k].reshape(-1) for k in keys] + [out["fine_grads"][k].reshape(-1) for k in keys])


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch_nerf_b200.parallel as par

    torch.set_num_threads(1)
    r, _, w = par.init_distributed("gloo")
    assert (r, w) == (rank, world)
    orc, pc, pf, focal, c2w, pix, target, u = _problem()
    a, b = par.shard_range(len(pix), rank, world)
    pix_s, tgt_s = par.shard_rays(torch.from_numpy(pix), torch.from_numpy(target), rank, world)
    assert pix_s.shape[0] == b - a
    flat = torch.from_numpy(_flat_grads(orc, pc, pf, focal, c2w, pix_s.numpy(), tgt_s.numpy(), [x[a:b] for x in u]))
    par.allreduce_mean_(flat)
    if rank == 0:
        ret["flat"] = flat.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import torch_nerf_b200.parallel as par

    for n in (0, 1, 7, 4096, 32768, 640000):
        for world in (1, 2, 3, 4, 8):
            edges = [par.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_single_process():
    sys.path.insert(0, ROOT)
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        dist_flat = np.array(ret["flat"])
    single = _flat_grads(*_problem())
    scale = np.abs(single).max()
    np.testing.assert_allclose(dist_flat / scale, single / scale, rtol=0, atol=2e-5)


def _sync_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch_nerf_b200.parallel as par

    torch.set_num_threads(1)
    par.init_distributed("gloo")
    torch.manual_seed(100 + rank)  # ranks seeded DIFFERENTLY on purpose
    flat = torch.nn.Parameter(torch.randn(1000))
    opt = torch.optim.Adam([flat], lr=5e-4 * (rank + 1), eps=1e-8)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.999)
    if rank == 0:  # only rank 0 has taken steps (e.g. it alone found the checkpoint)
        for _ in range(3):
            flat.grad = torch.randn(1000)
            opt.step()
            sched.step()
    par.broadcast_replica_state(flat, opt, sched)
    st = opt.state[flat]
    ret[rank] = (flat.detach().numpy().copy(), float(st["step"]), st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy(),
                 opt.param_groups[0]["lr"], sched.last_epoch)
    # one more identical step on both ranks must keep them identical
    flat.grad = torch.full((1000,), 0.25)
    opt.step()
    sched.step()
    ret[f"after{rank}"] = (flat.detach().numpy().copy(), opt.param_groups[0]["lr"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_replica_state_broadcast_makes_differently_seeded_ranks_identical():
    """Trainer(world > 1) must not depend on the caller seeding every rank identically."""
    sys.path.insert(0, ROOT)
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_sync_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
        a0, a1 = ret["after0"], ret["after1"]
    assert np.array_equal(r0[0], r1[0]) and r0[1] == r1[1] == 3.0
    assert np.array_equal(r0[2], r1[2]) and np.array_equal(r0[3], r1[3])
    assert r0[4] == r1[4] and r0[5] == r1[5] == 3
    assert np.array_equal(a0[0], a1[0]) and a0[1] == a1[1]
