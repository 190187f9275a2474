"""world_size-2 (and 3) gloo runs on the CPU of the multi-GPU host logic: rays dealt in
32-aligned tiles, traced per rank, gathered and put back in order.  The per-rank "trace" is the
CPU oracle, so this also checks the property BASELINE config 5 asks for: outputs identical for
every number of ranks."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grace-devel_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _dist
    import oracle as orc
    from util import clustered_spheres, isotropic_rays
    torch.set_num_threads(1)
    # rank 0 owns the particles; everybody else receives them (NCCL broadcast on GPUs)
    n = 6000
    s = torch.from_numpy(clustered_spheres(n, seed=3, n_halos=3)) if rank == 0 else torch.empty((n, 4))
    dist.broadcast(s, src=0)
    hs, _, _ = orc.sort_spheres(s.numpy(), 30)          # deterministic build on every rank
    tree = orc.build_tree(hs, orc.deltas_euclid(hs), 16)
    rays = torch.from_numpy(isotropic_rays(1504 * 2, seed=9))      # not a multiple of the tile
    out = _dist.sharded_trace(lambda r: torch.from_numpy(orc.trace_cumulative(r.numpy(), hs, tree)),
                              rays, torch.float32, tile=256)
    cnt = _dist.sharded_trace(lambda r: torch.from_numpy(orc.trace_hitcounts(r.numpy(), hs, tree)),
                              rays, torch.int32, tile=256)
    if rank == 0:
        np.savez(path, cum=out.numpy(), cnt=cnt.numpy(), nodes=tree.nodes)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, port):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.npz")
        mp.spawn(_worker, args=(world, port, path), nprocs=world, join=True)
        g = np.load(path)
        return g["cum"], g["cnt"], g["nodes"]


def test_tiles_partition_rays():
    import _dist
    for n, world, tile in ((32, 2, 32), (4096 * 5 + 64, 4, 4096), (3008, 3, 256), (64, 8, 32)):
        seen = np.zeros(n, int)
        for r in range(world):
            for a, b in _dist.tiles_of_rank(n, r, world, tile):
                assert a % 32 == 0 and (b - a) % 32 == 0
                seen[a:b] += 1
        assert (seen == 1).all()
    with pytest.raises(ValueError):
        _dist.tiles_of_rank(33, 0, 2)


def test_sharded_trace_identical_for_every_world_size():
    import oracle as orc
    from util import clustered_spheres, isotropic_rays
    hs, _, _ = orc.sort_spheres(clustered_spheres(6000, seed=3, n_halos=3), 30)
    tree = orc.build_tree(hs, orc.deltas_euclid(hs), 16)
    rays = isotropic_rays(1504 * 2, seed=9)
    ref_cum = orc.trace_cumulative(rays, hs, tree)
    ref_cnt = orc.trace_hitcounts(rays, hs, tree)
    for world, port in ((2, 29611), (3, 29612)):
        cum, cnt, nodes = _run(world, port)
        assert np.array_equal(nodes, tree.nodes)
        assert np.array_equal(cnt, ref_cnt)
        assert np.array_equal(cum.view(np.uint32), ref_cum.view(np.uint32))
