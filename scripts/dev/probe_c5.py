import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
pre = len(sys.argv) > 1 and sys.argv[1] == "pre"
if pre:
    s = gb.synth_gadget_spheres(1 << 24, 1234); t = gb.Tree(1 << 24, 32); gb.build_tree(s, t); print("pre-build 2^24 leaves", t.n_leaves); del s, t
n5 = 1 << 27
src = gb.synth_gadget_spheres(n5, 1234)
lo, hi = gb.min_max_x(src); c = (lo + hi) / 2
rays = gb.healpix_rays(None, 2048, 0, 1 << 15, c, c, c, 2.0 * (hi - lo))
for k in range(3):
    s5 = src.clone()
    tree = gb.Tree(n5, 32)
    gb.build_tree(s5, tree, key_bits=63)
    cum = torch.empty(1 << 15, dtype=torch.float32, device="cuda")
    gb.trace_cumulative_sph(rays, s5, tree, cum)
    print(json.dumps({"build": k, "n_leaves": tree.n_leaves, "root": int(tree.root_index_ptr.item()), "mean_cum": float(cum.double().mean()), "err": gb.device_error()}), flush=True)
    del tree, s5
