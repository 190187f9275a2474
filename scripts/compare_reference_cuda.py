#!/usr/bin/env python
"""Times the reference's own CUDA build (oracle/_ref/ref_driver) and this repo on the SAME
inputs on the same GPU, and checks their outputs against each other at full size.
Writes one JSON line (also to gpurun_out/compare_reference_cuda.json).

    python scripts/compare_reference_cuda.py [log2_particles=24] [log2_rays=20] [iters=3] [lists_log2_rays=0]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import grace_devel_b200 as gb
import refrun

lp = int(sys.argv[1]) if len(sys.argv) > 1 else 24
lr = int(sys.argv[2]) if len(sys.argv) > 2 else 20
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, r = 1 << lp, 1 << lr
mpl = 32

s0 = gb.synth_gadget_spheres(n, 1234)
h_s0 = s0.cpu().numpy()
lo, hi = gb.min_max_x(s0)
c = (lo + hi) / 2
length = 2 * (hi - lo)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


# ---- ours ----
work = torch.empty_like(s0)
deltas = torch.empty(n + 1, dtype=torch.float32, device="cuda")
state = {}


def our_sort():
    work.copy_(s0)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record(); gb.morton_keys30_sort_sph(work); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


t_sort = np.mean([our_sort() for _ in range(iters + 1)][1:])
t_deltas = timed(lambda: gb.euclidean_deltas_sph(work, deltas), iters)


def our_build():
    tree = gb.Tree(n, mpl)
    a, b = ev(), ev()
    a.record(); gb.ALBVH_sph(work, deltas, tree); b.record(); torch.cuda.synchronize()
    state["tree"] = tree
    return a.elapsed_time(b)


t_build = np.mean([our_build() for _ in range(iters + 1)][1:])
tree = state["tree"]
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
t_gen = timed(lambda: gb.uniform_random_rays(rays, c, c, c, length, 1234), iters)
cnt = torch.empty(r, dtype=torch.int32, device="cuda")
cum = torch.empty(r, dtype=torch.float32, device="cuda")
t_hit = timed(lambda: gb.trace_hitcounts_sph(rays, work, tree, cnt), iters)
t_cum = timed(lambda: gb.trace_cumulative_sph(rays, work, tree, cum), iters)
ours = dict(ms_keys_sort=float(t_sort), ms_deltas=t_deltas, ms_albvh=float(t_build), ms_gen_rays=t_gen,
            ms_hitcounts=t_hit, ms_cumulative=t_cum, n_leaves=tree.n_leaves)

# ---- reference ----
t0 = time.time()
ref, info = refrun.run(h_s0, "gen:%d:1234:%.9g:%.9g:%.9g:%.9g" % (r, c, c, c, length), mpl, 30, iters=iters,
                       lists=False, timeout=3000)
wall = time.time() - t0

parity = dict(
    sorted_spheres=bool(np.array_equal(ref["spheres_sorted"].view(np.uint32), work.cpu().numpy().view(np.uint32))),
    leaves=bool(np.array_equal(ref["leaves"][:, :2], tree.leaves.cpu().numpy()[:, :2])),
    nodes=bool(np.array_equal(ref["nodes"], tree.nodes.cpu().numpy())),
    root=bool(ref["root"] == int(tree.root_index_ptr.item())),
    rays=bool(np.array_equal(ref["rays"].view(np.uint32), rays.cpu().numpy().view(np.uint32))),
    hitcounts=bool(np.array_equal(ref["hitcounts"], cnt.cpu().numpy())),
    cumulative_bit_exact=bool(np.array_equal(ref["cumulative"].view(np.uint32), cum.cpu().numpy().view(np.uint32))),
    cumulative_max_rel=float(np.max(np.abs(ref["cumulative"] - cum.cpu().numpy()) / np.maximum(np.abs(ref["cumulative"]), 1e-30))),
)
# the reference's BEST iteration per stage (its per-call allocations make single iterations noisy)
best = info.get("min_ms") or dict(keys_sort=info["ms_keys_sort"], deltas=info["ms_deltas"], albvh=info["ms_albvh"],
                                  gen_rays=info["ms_gen_rays"], hitcounts=info["ms_hitcounts"], cumulative=info["ms_cumulative"])
ref_build = best["keys_sort"] + best["deltas"] + best["albvh"]
line = dict(particles=n, rays=r, max_per_leaf=mpl, key_bits=30, iters=iters, ours=ours, reference_cuda=info,
            speedup_vs_reference_best_iteration=dict(
                trace_cumulative=best["cumulative"] / t_cum, trace_hitcounts=best["hitcounts"] / t_hit,
                build=ref_build / (t_sort + t_deltas + t_build), gen_rays=best["gen_rays"] / t_gen),
            mrays_per_s=dict(ours=r / t_cum / 1e3, reference_cuda=r / best["cumulative"] / 1e3),
            mparticles_per_s=dict(ours=n / (t_sort + t_deltas + t_build) / 1e3, reference_cuda=n / ref_build / 1e3),
            parity=parity, reference_wall_s=wall,
            note="reference = GRACE headers patched only for CUDA-12 API removals (oracle/patch_ref.py), its "
                 "own launch configuration (MAX_BLOCKS = 112), timed with CUDA events around its public API calls")
print(json.dumps(line))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "compare_reference_cuda_2p%d_2p%d.json" % (lp, lr)), "w").write(json.dumps(line, indent=1))
