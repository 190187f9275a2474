"""Small pass over every kernel family for compute-sanitizer (dev tool): build, all trace modes
with forced splitting, hit lists + segmented sort + scan, generators, Gadget loader."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import grace_devel_b200 as gb
from util import clustered_spheres, isotropic_rays
s = torch.from_numpy(clustered_spheres(20000, seed=3)).cuda()
for bits in (30, 63):
    d = s.clone(); tree = gb.Tree(len(d), 16); gb.build_tree(d, tree, key_bits=bits)
rays = torch.from_numpy(isotropic_rays(2048, seed=4)).cuda()
cnt = torch.empty(2048, dtype=torch.int32, device="cuda"); cum = torch.empty(2048, dtype=torch.float32, device="cuda")
for mode in ("packet", "packet_wide", "ray", "packet_ref"):
    gb.set_trace_mode(mode)
    for budget, dyn, res in ((2048, 0, 0), (50, 0, 0), (50, 1, 0), (50, 0, 1)):
        gb.set_trace_budget(budget, eager=budget < 100); gb.set_trace_dynamic(dyn); gb.set_trace_resume(res)
        gb.trace_hitcounts_sph(rays, d, tree, cnt); gb.trace_cumulative_sph(rays, d, tree, cum)
        off = torch.empty(2048, dtype=torch.int32, device="cuda")
        idx, integ, dist = gb.trace_sph(rays, d, tree, off)
        gb.sort_by_distance(dist, off, idx, integ)
        tau = torch.empty_like(integ); gb.exclusive_segmented_scan(off, integ, tau)
gb.set_trace_mode("packet"); gb.set_trace_budget(1024); gb.set_trace_dynamic(0); gb.set_trace_resume(0)
r2 = torch.empty((4096, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(r2, 0.5, 0.5, 0.5, 2.0, 7)
gb.orthographic_projection_rays(None, 64, 32, (0.5, 0.5, 2.0), (0.5, 0.5, 0.5), (0, 1, 0), 1.0, 3.0)
gb.healpix_rays(None, 16, 0, 3072, 0.5, 0.5, 0.5, 2.0)
with tempfile.TemporaryDirectory() as t:
    p = os.path.join(t, "a.gdt"); gb.write_gadget(p, s.cpu()); gb.read_gadget(p)
torch.cuda.synchronize()
print("sanitize probe done, err", gb.device_error(), int(cnt.sum()))
