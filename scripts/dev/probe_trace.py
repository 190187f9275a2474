"""Dev probe: where does the trace time go?  (not part of the product or the tests)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import grace_devel_b200 as gb

n = 1 << 24; r = 1 << 20
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
counts = torch.empty(r, dtype=torch.int32, device="cuda")
out = torch.empty(r, dtype=torch.float32, device="cuda")

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for mode in ("ray", "packet_ref", "packet"):
    gb.set_trace_mode(mode)
    print(mode, "hitcounts ms", timeit(lambda: gb.trace_hitcounts_sph(rays, s, tree, counts)),
          "cumulative ms", timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out)))
h = counts.cpu().numpy().astype(np.int64)
print("hits: mean %.0f max %d p50 %d p99 %d p99.9 %d p99.99 %d" % (h.mean(), h.max(), *np.percentile(h, [50, 99, 99.9, 99.99])))
hs = s[:, 3].cpu().numpy()
print("h: min %.3g p1 %.3g p50 %.3g p99 %.3g max %.3g" % (hs.min(), *np.percentile(hs, [1, 50, 99]), hs.max()))
lv = tree.leaves.cpu().numpy()
print("leaves", len(lv), "mean count", lv[:, 1].mean())
for bud in (0, 1024, 2048, 4096, 8192):
    gb.set_trace_budget(bud)
    print("budget", bud, "hitcounts ms", timeit(lambda: gb.trace_hitcounts_sph(rays, s, tree, counts)),
          "cumulative ms", timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out)), "err", gb.device_error())
gb.set_trace_budget(2048)
pp = gb.trace_packet_profile_sph(rays, s, tree)
per = pp.pop("per_packet")
print("packet profile", pp)
cyc = per[:, 0]
print("per-packet cycles: mean %.3g max %.3g p99 %.3g ; sum/1e9 %.3f" % (cyc.mean(), cyc.max(), np.percentile(cyc, 99), cyc.sum() / 1e9))
w = np.argsort(-cyc)[:8]
print("heaviest packets", w, "\n cycles", cyc[w], "\n phaseA(nodes)", per[w, 1], "\n phaseB(loads)", per[w, 2], "\n phaseC(tests)", per[w, 3])
print("all packets: A %.3g B %.3g C %.3g (sum cycles /1e9)" % (per[:, 1].sum() / 1e9, per[:, 2].sum() / 1e9, per[:, 3].sum() / 1e9))
print("ref packet stats", gb.trace_stats_sph(rays, s, tree))
gb.set_trace_mode("packet")
tests, steps = gb.trace_ray_cost_sph(rays, s, tree)
t = tests.cpu().numpy().astype(np.int64); st_ = steps.cpu().numpy().astype(np.int64)
print("tests/ray: mean %.0f max %d p50 %d p99 %d p99.9 %d p99.99 %d" % (t.mean(), t.max(), *np.percentile(t, [50, 99, 99.9, 99.99])))
print("node steps/ray: mean %.0f max %d p50 %d p99 %d p99.9 %d p99.99 %d" % (st_.mean(), st_.max(), *np.percentile(st_, [50, 99, 99.9, 99.99])))
w = np.argsort(-st_)[:5]
print("worst rays", w, "steps", st_[w], "tests", t[w], "hits", h[w])
print(rays[torch.from_numpy(w).cuda()].cpu().numpy())
nd = tree.nodes.cpu().numpy(); fn = nd.view(np.float32)
# depth of the tree
L = len(lv); nn = L - 1
depth = np.zeros(nn + L, np.int32)
root = int(tree.root_index_ptr.item())
# iterative BFS by levels
frontier = np.array([root]); d = 0; maxd = 0
while len(frontier):
    depth[frontier] = d
    inner = frontier[frontier < nn]
    frontier = np.concatenate([nd[inner, 0], nd[inner, 1]]) if len(inner) else np.array([], np.int64)
    d += 1
print("tree depth", d, "mean leaf depth", depth[nn:].mean())
# box volume stats: ratio of node box extent to root
ext = np.stack([fn[:,5]-fn[:,4], fn[:,7]-fn[:,6], fn[:,13]-fn[:,12]],1)
print("left-child box extent percentiles (max axis):", np.percentile(ext.max(1), [50, 90, 99, 99.9, 100]))
