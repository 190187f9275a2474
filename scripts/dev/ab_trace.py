"""Dev A/B probe (not product, not a test): time the trace entry points of whichever
libgrace_b200 build GRACE_B200_LIB selects, at the bench workload, and print an output hash."""
import sys, os, hashlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb

lg_n = int(os.environ.get("AB_LOG2_N", "24")); lg_r = int(os.environ.get("AB_LOG2_R", "20"))
n = 1 << lg_n; r = 1 << lg_r
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
counts = torch.empty(r, dtype=torch.int32, device="cuda")
out = torch.empty(r, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for k in range(reps):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)

mode = os.environ.get("AB_MODE", "packet")
gb.set_trace_mode(mode)
if os.environ.get("AB_BUDGET"): gb.set_trace_budget(int(os.environ["AB_BUDGET"]))
if os.environ.get("AB_DYNAMIC"): gb.set_trace_dynamic(int(os.environ["AB_DYNAMIC"]))
if os.environ.get("AB_RESUME"): gb.set_trace_resume(int(os.environ["AB_RESUME"]))
res = {"lib": os.environ.get("GRACE_B200_LIB", "default").split("_")[-1], "mode": mode, "budget": os.environ.get("AB_BUDGET"), "dyn": os.environ.get("AB_DYNAMIC"), "resume": os.environ.get("AB_RESUME"), "lr": lg_r}
res["hitcounts_ms"] = timeit(lambda: gb.trace_hitcounts_sph(rays, s, tree, counts))
res["cumulative_ms"] = timeit(lambda: gb.trace_cumulative_sph(rays, s, tree, out))
res["counts_sha"] = hashlib.sha1(counts.cpu().numpy().tobytes()).hexdigest()[:12]
res["cum_sha"] = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12]
res["err"] = gb.device_error()
print(json.dumps(res))
