#!/usr/bin/env python
"""Hit lists (trace_sph + sort_by_distance): this repo vs the reference's own CUDA build on the
same GPU and inputs.   python scripts/compare_lists.py [log2_particles=24] [log2_rays=16] [iters=3]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import grace_devel_b200 as gb
import refrun
lp = int(sys.argv[1]) if len(sys.argv) > 1 else 24
lr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, r = 1 << lp, 1 << lr
s0 = gb.synth_gadget_spheres(n, 1234); h_s0 = s0.cpu().numpy()
lo, hi = gb.min_max_x(s0); c = (lo + hi) / 2; length = 2 * (hi - lo)
s = s0.clone(); tree = gb.Tree(n, 32); gb.build_tree(s, tree)
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, length, 1234)
off = torch.empty(r, dtype=torch.int32, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
t_tr = t_so = 0.0
idx = integ = dist = None
for k in range(iters + 1):
    del idx, integ, dist          # let the caching allocator reuse the blocks: no cudaMalloc inside the timed call
    a, b, d = ev(), ev(), ev()
    a.record(); idx, integ, dist = gb.trace_sph(rays, s, tree, off)
    b.record(); gb.sort_by_distance(dist, off, idx, integ)
    d.record(); torch.cuda.synchronize()
    if k: t_tr += a.elapsed_time(b) / iters; t_so += b.elapsed_time(d) / iters
ref, info = refrun.run(h_s0, "gen:%d:1234:%.9g:%.9g:%.9g:%.9g" % (r, c, c, c, length), 32, 30, iters=iters, lists=True, timeout=3000)
same = bool(np.array_equal(ref["offsets"], off.cpu().numpy()) and
            np.array_equal(ref["hit_dist"].view(np.uint32), dist.cpu().numpy().view(np.uint32)))
hits = idx.numel()
print(json.dumps(dict(particles=n, rays=r, hits=hits, ours=dict(ms_trace_lists=t_tr, ms_sort_by_distance=t_so),
                      reference_cuda=dict(ms_trace_lists=info["ms_trace_lists"], ms_sort_by_distance=info["ms_sort_by_distance"]),
                      speedup=dict(trace_lists=info["ms_trace_lists"] / t_tr, sort_by_distance=info["ms_sort_by_distance"] / t_so,
                                   both=(info["ms_trace_lists"] + info["ms_sort_by_distance"]) / (t_tr + t_so)),
                      mhits_per_s=dict(ours=hits / (t_tr + t_so) / 1e3, reference_cuda=hits / (info["ms_trace_lists"] + info["ms_sort_by_distance"]) / 1e3),
                      offsets_and_sorted_distances_bit_identical=same)))
