#!/usr/bin/env python
"""BASELINE config 5 on N GPUs of one box: 2^27 particles (63-bit keys), 2^24 HEALPix NESTED rays
from the box centre; particles broadcast with NCCL, tree built on every rank (deterministic),
rays dealt round-robin in 4096-ray tiles, per-ray column densities gathered with one all_gather
and put back in ray order (strong scaling: the total work is fixed).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config5_mgpu.py [--log2-n 27] [--log2-rays 24]
Prints one JSON line on rank 0.  With --check rank 0 also traces all rays alone and compares bits."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--log2-n", type=int, default=27)
ap.add_argument("--log2-rays", type=int, default=24)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import grace_devel_b200 as gb

n, r = 1 << args.log2_n, 1 << args.log2_rays
dev = torch.device("cuda", local)
def ev(): return torch.cuda.Event(enable_timing=True)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

s = gb.synth_gadget_spheres(n, 1234) if rank == 0 else torch.empty((n, 4), dtype=torch.float32, device=dev)
barrier(); a, b = ev(), ev(); a.record()
if world > 1: dist.broadcast(s, src=0)
b.record(); barrier(); t_bcast = a.elapsed_time(b)
tree = gb.Tree(n, 32)
a, b = ev(), ev(); a.record(); gb.build_tree(s, tree, key_bits=63); b.record(); barrier(); t_build = a.elapsed_time(b)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
nside = 1 << ((args.log2_rays - 2) // 2 + (0 if (args.log2_rays - 2) % 2 == 0 else 1))   # 12*nside^2 >= r
while 12 * nside * nside < r: nside *= 2
rays = gb.healpix_rays(None, nside, 0, r, c, c, c, 2 * (hi - lo))

def trace(local_rays):
    out = torch.empty(local_rays.shape[0], dtype=torch.float32, device=dev)
    gb.trace_cumulative_sph(local_rays, s, tree, out)
    return out

full = gb.sharded_trace(trace, rays, torch.float32)     # warm-up
times = []
for k in range(args.steps):
    barrier(); a, b = ev(), ev(); a.record()
    full = gb.sharded_trace(trace, rays, torch.float32)
    b.record(); barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = sum(times) / len(times)
line = dict(config="one_to_many_rays (HEALPix NESTED pixels [0, 2^%d), nside %d), 2^%d particles, 63-bit keys" % (args.log2_rays, nside, args.log2_n),
            n_gpus=world, particles=n, rays=r, broadcast_ms=t_bcast, build_ms_per_rank=t_build, trace_gather_ms=ms,
            mrays_s=r / ms / 1e3, scaling="strong", device_error=gb.device_error())
if args.check and rank == 0:
    alone = trace(rays)
    line["identical_to_single_gpu"] = bool(torch.equal(alone.view(torch.int32), full.view(torch.int32)))
if rank == 0: print(json.dumps(line))
if world > 1: dist.destroy_process_group()
