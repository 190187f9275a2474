"""Dev A/B (round 2): trace times at several ray counts, with result hashes (not part of the product)."""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb

n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
lrs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [23, 20, 17, 15, 12]
tag = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("GRACE_B200_LIB", "default")
if os.environ.get("AB_BUDGET"):
    gb.set_trace_budget(int(os.environ["AB_BUDGET"]))
    tag += "_b" + os.environ["AB_BUDGET"]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


for lr in lrs:
    r = 1 << lr
    rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
    counts = torch.empty(r, dtype=torch.int32, device="cuda")
    out = torch.empty(r, dtype=torch.float32, device="cuda")
    hc = timeit(lambda: gb.trace_hitcounts_sph(rays, s, tree, counts))
    st_c = gb.trace_balance_stats()
    cu = timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out))
    st_u = gb.trace_balance_stats()
    err = gb.device_error()
    print(json.dumps({"tag": tag, "lr": lr, "hitcounts_ms": round(hc[0], 3), "hitcounts_mean": round(hc[1], 3),
                      "cumulative_ms": round(cu[0], 3), "cumulative_mean": round(cu[1], 3),
                      "counts_sha": hashlib.sha1(counts.cpu().numpy().tobytes()).hexdigest()[:12],
                      "cum_sha": hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12], "err": err, "lb_count": st_c, "lb_cum": st_u}), flush=True)
