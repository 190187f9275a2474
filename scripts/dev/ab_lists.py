"""Dev A/B: trace_sph (count + fill) and sort at 2^24 particles: isotropic 2^17 rays and config-4 tiles; one-pass on/off."""
import sys, os, json, hashlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
def ev(): return torch.cuda.Event(enable_timing=True)
def run(rays, tag):
    r = rays.shape[0]
    off = torch.empty(r, dtype=torch.int32, device="cuda")
    idx = integ = dist = None
    ts = []
    for k in range(4):
        del idx, integ, dist
        a, b = ev(), ev()
        a.record(); idx, integ, dist = gb.trace_sph(rays, s, tree, off); b.record(); torch.cuda.synchronize()
        if k: ts.append(a.elapsed_time(b))
    h = hashlib.sha1(idx.cpu().numpy().tobytes() + dist.cpu().numpy().tobytes() + integ.cpu().numpy().tobytes() + off.cpu().numpy().tobytes()).hexdigest()[:12]
    print(json.dumps({"tag": tag, "rays": r, "hits": idx.numel(), "trace_sph_ms": round(min(ts), 3), "sha": h, "err": gb.device_error()}), flush=True)
for lr in (17, 14):
    rays = torch.empty((1 << lr, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
    run(rays, "iso_2p%d" % lr)
mins = [float(v) for v in gb.min_vec4(s).cpu()]; maxs = [float(v) for v in gb.max_vec4(s).cpu()]
cx, cy, cz = [(mins[k] + maxs[k]) / 2 for k in range(3)]
span = [maxs[k] - mins[k] for k in range(3)]; span[0] = span[1] = max(span[0], span[1])
img = gb.orthographic_projection_rays(None, 4096, 4096, (cx, cy, span[2]), (cx, cy, cz), (0, 1, 0), span[1], 2 * span[2])
run(img[100 * 65536: 101 * 65536].contiguous(), "ortho_tile")
t0 = time.perf_counter()
tot = gb.trace_sorted_tiles(img[:32 * 65536].contiguous(), s, tree, 1 << 28, lambda *a: None, 65536)
torch.cuda.synchronize(); t0 = time.perf_counter()
tot = gb.trace_sorted_tiles(img[:32 * 65536].contiguous(), s, tree, 1 << 28, lambda *a: None, 65536)
torch.cuda.synchronize()
print(json.dumps({"tag": "tiles_32", "hits": tot, "ms_per_tile": round((time.perf_counter() - t0) * 1e3 / 32, 3)}), flush=True)
