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
