#!/usr/bin/env python
"""Measures every BASELINE.json configuration on one B200 and prints one JSON object
(committed as profiles/<round>_configs.json).  Not the driver's bench (bench.py is): the
numbers here are the per-config table of DESIGN.md section 7.

  python scripts/bench_configs.py [--configs 1,2,3,4,5] [--log2-n5 27]
"""
import argparse, json, os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import grace_devel_b200 as gb

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,2,3,4,5")
ap.add_argument("--log2-n", type=int, default=24)
ap.add_argument("--log2-n5", type=int, default=27)
ap.add_argument("--image", type=int, default=4096)
ap.add_argument("--tile-rays", type=int, default=1 << 16)
args = ap.parse_args()
want = set(int(c) for c in args.configs.split(","))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak}

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for k in range(reps):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.mean(ts)

def build_stages(n, bits):
    s0 = gb.synth_gadget_spheres(n, 1234)
    acc = {"bounds_keys_sort": [], "deltas": [], "albvh": []}
    for k in range(4):
        s = s0.clone(); tree = gb.Tree(n, 32); deltas = torch.empty(n + 1, dtype=torch.float32, device="cuda")
        flush.fill_(k); e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); (gb.morton_keys30_sort_sph if bits == 30 else gb.morton_keys63_sort_sph)(s)
        e[1].record(); gb.euclidean_deltas_sph(s, deltas)
        e[2].record(); gb.ALBVH_sph(s, deltas, tree)
        e[3].record(); torch.cuda.synchronize()
        if k:
            acc["bounds_keys_sort"].append(e[0].elapsed_time(e[1])); acc["deltas"].append(e[1].elapsed_time(e[2]))
            acc["albvh"].append(e[2].elapsed_time(e[3]))
    ms = {k: statistics.mean(v) for k, v in acc.items()}
    total = sum(ms.values()); L = tree.n_leaves
    bpp = (116 if bits == 30 else 128) + 88.0 * L / n
    return s, tree, dict(ms=ms, total_ms=total, mparticles_s=n / total / 1e3, n_leaves=L,
                         algorithmic_bytes_per_particle=bpp, hbm_frac=bpp * n / (total * 1e-3) / 1e9 / peak)

if 1 in want:      # hitcounts: 2^16 uniform spheres, 2^14 isotropic rays, vs host brute force
    import oracle
    from util import uniform_spheres
    s = uniform_spheres(1 << 16, seed=1, rmax=0.1)
    d_s = torch.from_numpy(s).cuda(); tree = gb.Tree(len(s), 32)
    gb.build_tree(d_s, tree, (0, 0, 0), (1, 1, 1))
    rays = torch.empty((1 << 14, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, 0.5, 0.5, 0.5, 2.0, 1234)
    cnt = torch.empty(1 << 14, dtype=torch.int32, device="cuda")
    ms = timed(lambda: gb.trace_hitcounts_sph(rays, d_s, tree, cnt), reps=5)
    t0 = time.perf_counter(); ref = oracle.brute_hitcounts(rays.cpu().numpy(), d_s.cpu().numpy()); cpu_s = time.perf_counter() - t0
    out["config1_hitcounts"] = dict(ms=ms, mrays_s=(1 << 14) / ms / 1e3, equals_host_brute_force=bool(np.array_equal(ref, cnt.cpu().numpy())),
                                    total_hits=int(ref.sum()), cpu_brute_force_s=cpu_s, cpu_threads=oracle.num_threads())

if want & {2, 3, 4}:
    n = 1 << args.log2_n
    s, tree, b30 = build_stages(n, 30)
    if 2 in want:
        _, _, b63 = build_stages(n, 63)
        out["config2_profile_tree_gadget"] = {"n": n, "keys30": b30, "keys63": b63}

if 3 in want:
    r = 1 << 20
    lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
    rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
    t_gen = timed(lambda: gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), 1234))
    cnt = torch.empty(r, dtype=torch.int32, device="cuda"); cum = torch.empty(r, dtype=torch.float32, device="cuda")
    t_cnt = timed(lambda: gb.trace_hitcounts_sph(rays, s, tree, cnt)); t_cum = timed(lambda: gb.trace_cumulative_sph(rays, s, tree, cum))
    out["config3_profile_trace_gadget"] = dict(n=n, rays=r, gen_rays_ms=t_gen, hitcounts_ms=t_cnt, cumulative_ms=t_cum,
        mrays_s_cumulative=r / t_cum / 1e3, mrays_s_hitcounts=r / t_cnt / 1e3, hits_per_ray=float(cnt.sum().item()) / r)

if 4 in want:
    side = args.image; r = side * side
    mins = [float(v) for v in gb.min_vec4(s).cpu()]; maxs = [float(v) for v in gb.max_vec4(s).cpu()]
    cx, cy, cz = [(mins[k] + maxs[k]) / 2 for k in range(3)]
    span = [maxs[k] - mins[k] for k in range(3)]; span[0] = span[1] = max(span[0], span[1])
    rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
    # tests/helper/rays.cuh:55-79 orthogonal_rays_z (mins.w = maxs.w = 0 as in project_gadget.cu:70-74)
    t_gen = timed(lambda: gb.orthographic_projection_rays(rays, side, side, (cx, cy, span[2]), (cx, cy, cz), (0, 1, 0), span[1], 2 * span[2]))
    cum = torch.empty(r, dtype=torch.float32, device="cuda")
    t_cum = timed(lambda: gb.trace_cumulative_sph(rays, s, tree, cum), reps=2)
    area = (span[0] / side) * (span[1] / side)
    mass = float(cum.double().sum().item()) * area       # each particle's kernel integrates to 1 over the plane
    # sorted hit lists of the WHOLE image, streamed in ray tiles by the library (count -> scan -> fill -> sort ->
    # consumer per tile, reused buffers, sort of tile k overlapping the count of tile k + 1)
    tile = args.tile_rays
    seen = dict(tiles=0, hits=0, bad=0)
    scan_info = {}

    def consume(first, off, idx, integ, dist):
        seen["tiles"] += 1
        seen["hits"] += dist.numel()
        if seen["tiles"] == 2 and dist.numel():   # the step after the path: optical-depth style scan along the sorted lists
            tau = torch.empty_like(integ)
            t_scan = timed(lambda: gb.exclusive_segmented_scan(off, integ, tau), reps=3)
            scan_info.update(hits=idx.numel(), ms=t_scan, ghits_s=idx.numel() / t_scan / 1e6,
                             hbm_frac=8.0 * idx.numel() / (t_scan * 1e-3) / 1e9 / peak)
        if seen["tiles"] % 64 == 1 and dist.numel():   # sortedness of a sample of tiles
            n_hits = dist.numel()
            boundary = torch.zeros(n_hits, dtype=torch.bool, device=off.device)
            boundary[off[off < n_hits].long()] = True
            seen["bad"] += int(((dist[1:] < dist[:-1]) & ~boundary[1:]).sum())

    gb.trace_sorted_tiles(rays[: 4 * tile], s, tree, 1 << 28, lambda *a: None, tile)      # buffers allocated, kernels loaded
    t0 = time.perf_counter()
    total_hits = gb.trace_sorted_tiles(rays, s, tree, 1 << 28, consume, tile)
    torch.cuda.synchronize()
    t_lists_all = (time.perf_counter() - t0) * 1e3
    ok_sorted = seen["bad"] == 0
    n_tiles = seen["tiles"]; hits = total_hits; t_lists = t_lists_all; t_sort = 0.0; rel = None
    out["config4_project_gadget"] = dict(n=n, image=[side, side], gen_rays_ms=t_gen, cumulative_ms=t_cum, mrays_s_cumulative=r / t_cum / 1e3,
        mass_recovered_over_n=mass / n, lists=dict(tiles=n_tiles, rays_per_tile=tile, hits=hits, trace_ms=t_lists, sort_ms=t_sort,
        whole_image=True, wall_ms=t_lists_all, mrays_s=r / t_lists_all / 1e3, mhits_s=hits / t_lists_all / 1e3, sorted=bool(ok_sorted),
        list_sum_vs_cumulative_max_rel=rel), exclusive_segmented_scan=scan_info)
    del rays, cum

if 6 in want:     # N1: Gadget-2 snapshot -> device float4 records (the step before the path)
    import subprocess, tempfile
    n6 = 1 << args.log2_n
    h6 = gb.synth_gadget_spheres(n6, 1234).cpu()
    d = tempfile.mkdtemp(dir="/tmp")
    path = os.path.join(d, "snap.gdt")
    gb.write_gadget(path, h6)
    dst = torch.empty((n6, 4), dtype=torch.float32, device="cuda")
    ts = []
    for k in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gb.read_gadget(path, dst); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ok = bool(torch.equal(dst.cpu().view(torch.int32), h6.view(torch.int32)))
    info = dict(n=n6, file_bytes=os.path.getsize(path), bytes_read=16 * n6, seconds_first=ts[0], seconds=min(ts[1:]),
                gb_s=16 * n6 / min(ts[1:]) / 1e9, equals_input=ok,
                note="file in the page cache; wall clock from the call to the last record on the device")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_gadget_driver")
    if os.path.exists(ref):
        r = subprocess.run([ref, path, os.path.join(d, "ref.bin")], capture_output=True, text=True, timeout=900)
        if r.returncode == 0:
            info["reference_reader_seconds"] = json.loads(r.stdout.strip().splitlines()[-1])["seconds_file_to_device"]
    import shutil; shutil.rmtree(d, ignore_errors=True)
    out["n1_read_gadget"] = info
    del dst, h6

if 5 in want:
    del_s = None
    if want & {2, 3, 4}: del s, tree
    torch.cuda.empty_cache()
    n5 = 1 << args.log2_n5
    s5 = gb.synth_gadget_spheres(n5, 1234)
    tree5 = gb.Tree(n5, 32)
    t0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0[0].record(); gb.build_tree(s5, tree5, key_bits=63); t0[1].record(); torch.cuda.synchronize()
    build_ms = t0[0].elapsed_time(t0[1])
    lo, hi = gb.min_max_x(s5); c = (lo + hi) / 2; length = 2 * (hi - lo)
    r5 = 1 << 24
    rays = torch.empty((r5, 7), dtype=torch.float32, device="cuda")
    t_gen = timed(lambda: gb.healpix_rays(rays, 2048, 0, r5, c, c, c, length), reps=2)
    cum = torch.empty(r5, dtype=torch.float32, device="cuda")
    t_cum = timed(lambda: gb.trace_cumulative_sph(rays, s5, tree5, cum), reps=2)
    out["config5_one_to_many_healpix"] = dict(n=n5, rays=r5, nside=2048, build_ms_incl_first_call=build_ms, n_leaves=tree5.n_leaves,
        gen_rays_ms=t_gen, cumulative_ms=t_cum, mrays_s=r5 / t_cum / 1e3, device_error=gb.device_error(),
        mean_column_density=float(cum.double().mean().item()))
    # run (B) of SURVEY 8d: the full sky at nside = 1024 (12 * 1024^2 = 12 582 912 rays, a multiple of 32)
    rb = 12 * 1024 * 1024
    t_gen_b = timed(lambda: gb.healpix_rays(rays[:rb], 1024, 0, rb, c, c, c, length), reps=2)
    t_cum_b = timed(lambda: gb.trace_cumulative_sph(rays[:rb], s5, tree5, cum[:rb]), reps=2)
    out["config5b_full_sky_healpix"] = dict(n=n5, rays=rb, nside=1024, gen_rays_ms=t_gen_b, cumulative_ms=t_cum_b,
        mrays_s=rb / t_cum_b / 1e3, device_error=gb.device_error(),
        mean_column_density=float(cum[:rb].double().mean().item()))
print(json.dumps(out))
