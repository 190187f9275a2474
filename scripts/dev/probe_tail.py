"""Dev probe (round 2): anatomy of the trace tail -- what do the heaviest packets look like?
(not part of the product or the tests)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import grace_devel_b200 as gb

n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for lr in (23, 20, 17):
    r = 1 << lr
    rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
    counts = torch.empty(r, dtype=torch.int32, device="cuda")
    out = torch.empty(r, dtype=torch.float32, device="cuda")
    res = {"lr": lr,
           "hitcounts_ms": timeit(lambda: gb.trace_hitcounts_sph(rays, s, tree, counts)),
           "cumulative_ms": timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out))}
    if lr <= 20:
        gb.set_trace_budget(0)
        res["cumulative_ms_nosplit"] = timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out), 1)
        gb.set_trace_budget(1024)
    print(json.dumps(res), flush=True)
    if lr > 20:
        continue
    h = counts.cpu().numpy().astype(np.int64)
    print(" hits/ray: mean %.0f max %d p50 %d p99 %d p99.9 %d p99.99 %d" % (h.mean(), h.max(), *np.percentile(h, [50, 99, 99.9, 99.99])))
    ph = h.reshape(-1, 32)
    print(" hits/packet: mean %.0f max %d ; packet max-lane/mean-lane of the 8 heaviest: %s" % (
        ph.sum(1).mean(), ph.sum(1).max(),
        [round(float(ph[i].max() / max(ph[i].mean(), 1)), 2) for i in np.argsort(-ph.sum(1))[:8]]))
    pp = gb.trace_packet_profile_sph(rays, s, tree)
    per = pp.pop("per_packet")
    print(" packet profile", pp)
    cyc = per[:, 0]
    print(" per-packet cycles: mean %.3g max %.3g p99 %.3g p99.9 %.3g; sum/1e9 %.3f" % (
        cyc.mean(), cyc.max(), np.percentile(cyc, 99), np.percentile(cyc, 99.9), cyc.sum() / 1e9))
    w = np.argsort(-cyc)[:12]
    print(" heaviest packets", w.tolist())
    print("  cycles     ", cyc[w].tolist())
    print("  node steps ", per[w, 1].tolist())
    print("  leaf visits", per[w, 2].tolist())
    print("  kept       ", per[w, 3].tolist())
    print("  hits       ", ph.sum(1)[w].tolist())
    # cumulative share of cycles held by the heaviest x % of the packets
    srt = np.sort(cyc)[::-1]
    cs = np.cumsum(srt) / srt.sum()
    for frac in (0.001, 0.01, 0.05, 0.1):
        print("  top %.1f%% of packets hold %.1f%% of the cycles" % (100 * frac, 100 * cs[int(frac * len(srt))]))
    # ideal: sum of cycles / warp slots vs the longest packet
    slots = 148 * 7 * 4
    print("  sum cycles / %d slots = %.3g cycles = %.2f ms at 1.9 GHz; longest packet %.2f ms" % (
        slots, cyc.sum() / slots, cyc.sum() / slots / 1.9e6, cyc.max() / 1.9e6))
