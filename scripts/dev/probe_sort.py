"""Dev probe: segment-length distribution and sort_by_distance time on orthographic tiles (config 4)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234); tree = gb.Tree(n, 32); gb.build_tree(s, tree)
mins = [float(v) for v in gb.min_vec4(s).cpu()]; maxs = [float(v) for v in gb.max_vec4(s).cpu()]
cx, cy, cz = [(mins[k] + maxs[k]) / 2 for k in range(3)]
span = [maxs[k] - mins[k] for k in range(3)]; span[0] = span[1] = max(span[0], span[1])
side = 4096
rays = gb.orthographic_projection_rays(None, side, side, (cx, cy, span[2]), (cx, cy, cz), (0, 1, 0), span[1], 2 * span[2])
tile = 1 << 16
tiles = [int(x) for x in os.environ.get("AB_TILES", "0,1,2,3,4,5,6,7").split(",")]
off = torch.empty(tile, dtype=torch.int32, device="cuda")
tot_ms = 0.0; hist = np.zeros(6, np.int64); elems = np.zeros(6, np.int64); longest = 0
for k in tiles:
    sub = rays[(k * (side * side // 8)) // 32 * 32:][:tile].contiguous()
    idx, integ, dist = gb.trace_sph(sub, s, tree, off)
    o = off.cpu().numpy().astype(np.int64); lens = np.diff(np.append(o, idx.numel()))
    for c, (lo, hi) in enumerate([(0, 32), (33, 512), (513, 2048), (2049, 8192), (8193, 65536), (65537, 1 << 40)]):
        m = (lens >= lo) & (lens <= hi); hist[c] += m.sum(); elems[c] += lens[m].sum()
    longest = max(longest, int(lens.max()))
    d2, i2, g2 = dist.clone(), idx.clone(), integ.clone()
    gb.sort_by_distance(d2, off, i2, g2)                      # warm (workspace growth)
    d2.copy_(dist); i2.copy_(idx); g2.copy_(integ)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gb.sort_by_distance(d2, off, i2, g2); b.record(); torch.cuda.synchronize()
    tot_ms += a.elapsed_time(b)
    del idx, integ, dist, d2, i2, g2
print(json.dumps(dict(lib=os.environ.get("GRACE_B200_LIB", "default").split("_")[-1], tiles=len(tiles), sort_ms=tot_ms,
                      segments=hist.tolist(), elements=elems.tolist(), longest=longest)))
