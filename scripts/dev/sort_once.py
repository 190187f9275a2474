import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
rays = torch.empty((1 << 17, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
off = torch.empty(1 << 17, dtype=torch.int32, device="cuda")
idx, integ, dist = gb.trace_sph(rays, s, tree, off)
ts = []
for k in range(4):
    d2, i2, g2 = dist.clone(), idx.clone(), integ.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gb.sort_by_distance(d2, off, i2, g2); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
h = hashlib.sha1(d2.cpu().numpy().tobytes() + i2.cpu().numpy().tobytes() + g2.cpu().numpy().tobytes()).hexdigest()[:12]
print(json.dumps({"sort_ms": [round(t, 3) for t in ts], "hits": idx.numel(), "sha": h}))
