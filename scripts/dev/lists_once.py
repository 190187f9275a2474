import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
rays = torch.empty((1 << 17, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
if os.environ.get('POOL'): gb.set_trace_pool(int(os.environ['POOL']) << 20)
off = torch.empty(1 << 17, dtype=torch.int32, device="cuda")
for k in range(2):
    idx, integ, dist = gb.trace_sph(rays, s, tree, off); torch.cuda.synchronize()
    del idx, integ, dist
gb.sort_by_distance(*(lambda i, g, d: (d, off, i, g))(*gb.trace_sph(rays, s, tree, off))) if hasattr(gb, "sort_by_distance") else None
torch.cuda.synchronize()
