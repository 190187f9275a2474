"""Dev A/B: trace_sph on orthographic tiles (config 4) for several split budgets."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234); tree = gb.Tree(n, 32); gb.build_tree(s, tree)
mins = [float(v) for v in gb.min_vec4(s).cpu()]; maxs = [float(v) for v in gb.max_vec4(s).cpu()]
cx, cy, cz = [(mins[k] + maxs[k]) / 2 for k in range(3)]
span = [maxs[k] - mins[k] for k in range(3)]; span[0] = span[1] = max(span[0], span[1])
side = 4096
rays = gb.orthographic_projection_rays(None, side, side, (cx, cy, span[2]), (cx, cy, cz), (0, 1, 0), span[1], 2 * span[2])
tile = 1 << 16
subs = [rays[(k * (side * side // 8)) // 32 * 32:][:tile].contiguous() for k in range(8)]
off = torch.empty(tile, dtype=torch.int32, device="cuda")
cum = torch.empty(tile, dtype=torch.float32, device="cuda")
for b in [int(x) for x in os.environ.get("AB_BUDGETS", "512,1024,2048,4096").split(",")]:
    gb.set_trace_budget(b)
    t_l = t_c = 0.0
    for sub in subs:
        idx, integ, dist = gb.trace_sph(sub, s, tree, off); del idx, integ, dist
        a, e, f = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); idx, integ, dist = gb.trace_sph(sub, s, tree, off); e.record()
        gb.trace_cumulative_sph(sub, s, tree, cum); f.record(); torch.cuda.synchronize()
        t_l += a.elapsed_time(e); t_c += e.elapsed_time(f); del idx, integ, dist
    print(json.dumps(dict(budget=b, trace_sph_ms=t_l, cumulative_ms=t_c)))
