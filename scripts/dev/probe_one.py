"""Dev probe: N calls of one trace entry point at one ray count (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
lr = int(sys.argv[1]); what = sys.argv[2]; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
r = 1 << lr
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
counts = torch.empty(r, dtype=torch.int32, device="cuda")
out = torch.empty(r, dtype=torch.float32, device="cuda")
for _ in range(reps):
    if what == "count":
        gb.trace_hitcounts_sph(rays, s, tree, counts)
    else:
        gb.trace_cumulative_sph(rays, s, tree, out)
    torch.cuda.synchronize()
print("ok", gb.trace_balance_stats(), gb.device_error())
