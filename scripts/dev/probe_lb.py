"""Dev probe: work-stealing debug counters (needs the -DPK_DEBUG_LB build, GRACE_B200_LIB=...)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
n = 1 << 24
s = gb.synth_gadget_spheres(n, 1234)
tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
dbg = gb.lib.grace_b200_debug_lb
dbg.argtypes = [ctypes.c_void_p, ctypes.c_int]
for lr in [int(a) for a in sys.argv[1].split(",")]:
    r = 1 << lr
    rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234)
    counts = torch.empty(r, dtype=torch.int32, device="cuda")
    out = torch.empty(r, dtype=torch.float32, device="cuda")
    for what in ("count", "cum"):
        fn = (lambda: gb.trace_hitcounts_sph(rays, s, tree, counts)) if what == "count" else (lambda: gb.trace_cumulative_sph(rays, s, tree, out))
        fn(); torch.cuda.synchronize()
        dbg(None, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 32)()
        dbg(buf, 0)
        d = list(buf)
        t0 = d[7]
        print("lr %d %s: %.2f ms | task steps sum %d max %d (n<16 steps: %d) | packet steps sum %d max %d | first thief +%.2f ms, last packet +%.2f ms, last task +%.2f ms | thief time until a task: sum %.1f ms, until exit: sum %.1f ms | served %d, refused by victim %d | thief attempts: got %d, refused %d, CAS lost %d, empty scans %d | %s" % (
            lr, what, a.elapsed_time(b), d[0], d[1], d[14], d[2], d[3], (d[4] - t0) / 1e6, (d[5] - t0) / 1e6, (d[6] - t0) / 1e6 if d[6] else 0,
            d[8] / 1e6, d[9] / 1e6, d[10], d[11], d[16], d[17], d[18], d[19], gb.trace_balance_stats()), flush=True)
        if what == "cum":
            print("   fold: chunks %d blocks %d | records visited sum %d max/root %d | root fold time max %.2f ms sum %.1f ms | inline steps %d" % (
                d[20], d[21], d[22], d[23], d[24] / 1e6, d[25] / 1e6, d[26]), flush=True)
