"""Ray sharding for multi-GPU tracing: the tree and the sorted particles are
replicated, rays are dealt to the ranks in 32-aligned tiles (round-robin, so direction-sorted
rays of very different cost spread evenly), each rank traces its tiles, and the per-ray
results are gathered with one all_gather and put back in ray order.

Packets are 32 consecutive rays (include/grace/cuda/kernels/bintree_trace.cuh:75,231-238), so
tile boundaries are multiples of 32 and every ray sits in the same packet as in a
single-GPU run: outputs are identical for every number of ranks.

Backend-agnostic (NCCL on GPUs, gloo on CPU for the tests); the only collectives are the
caller's broadcast of the particles and the all_gather here."""
import torch
import torch.distributed as dist

TILE = 4096


def tiles_of_rank(n_rays, rank, world, tile=TILE):
    """[(start, stop), ...] of the tiles owned by `rank`."""
    if n_rays % 32:
        raise ValueError("Number of rays must be a multiple of the warp size (32).")
    if tile % 32:
        raise ValueError("tile must be a multiple of 32")
    n_tiles = (n_rays + tile - 1) // tile
    return [(t * tile, min((t + 1) * tile, n_rays)) for t in range(rank, n_tiles, world)]


def local_ray_count(n_rays, rank, world, tile=TILE):
    return sum(b - a for a, b in tiles_of_rank(n_rays, rank, world, tile))


def padded_local_count(n_rays, world, tile=TILE):
    """Rays per rank after padding every rank to the largest share (all_gather needs equal sizes)."""
    return max(local_ray_count(n_rays, r, world, tile) for r in range(world))


def take_local(rays, rank, world, tile=TILE):
    """The rank's rays, tiles concatenated in order, padded with copies of its last ray."""
    n = rays.shape[0]
    if n % (tile * world) == 0:       # every rank owns the same number of full tiles: one strided copy
        return rays.view(n // (tile * world), world, tile, *rays.shape[1:])[:, rank].reshape(
            n // world, *rays.shape[1:]).contiguous()
    parts = [rays[a:b] for a, b in tiles_of_rank(n, rank, world, tile)]
    local = torch.cat(parts, 0) if parts else rays[:0]
    pad = padded_local_count(n, world, tile) - local.shape[0]
    if pad:
        filler = (local[-1:] if local.shape[0] else rays[:1]).expand(pad, *rays.shape[1:])
        local = torch.cat([local, filler], 0)
    return local.contiguous()


def scatter_back(gathered, n_rays, world, tile=TILE):
    """gathered: [world * padded_local_count] results in rank-major order -> ray order."""
    per = padded_local_count(n_rays, world, tile)
    if n_rays % (tile * world) == 0:  # inverse of the strided copy in take_local
        rest = tuple(gathered.shape[1:])
        return gathered.view(world, n_rays // (tile * world), tile, *rest).transpose(0, 1).reshape(n_rays, *rest)
    out = torch.empty((n_rays,) + tuple(gathered.shape[1:]), dtype=gathered.dtype, device=gathered.device)
    for r in range(world):
        pos = r * per
        for a, b in tiles_of_rank(n_rays, r, world, tile):
            out[a:b] = gathered[pos:pos + (b - a)]
            pos += b - a
    return out


def sharded_trace(trace_fn, rays, out_dtype, tile=TILE, group=None):
    """Every rank holds the full `rays`; returns the full per-ray result on every rank.

    trace_fn(local_rays) -> tensor [len(local_rays)] (e.g. a closure over
    trace_cumulative_sph with the replicated spheres and tree)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays.shape[0]
    local = take_local(rays, rank, world, tile)
    local_out = trace_fn(local).to(out_dtype)
    if world == 1:
        return scatter_back(local_out, n, 1, tile)
    gathered = torch.empty(world * local_out.shape[0], dtype=out_dtype, device=local_out.device)
    dist.all_gather_into_tensor(gathered, local_out.contiguous(), group=group)
    return scatter_back(gathered, n, world, tile)
