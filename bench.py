#!/usr/bin/env python
"""bench.py -- GRACE hot path on B200: SPH cumulative column-density trace (Mrays/s).

Workload (BASELINE.json configs[2] "profile_trace_gadget", the configuration the metric
"SPH trace Mrays/s (2^24 particles, 1-8 B200)" is quoted on; it fits one GPU):
  2^24 synthetic Gadget-shaped SPH particles (float4 x,y,z,h), ALBVH with max_per_leaf=32,
  30-bit keys, Euclidean deltas (tests/helper/tree.cuh:15-27 recipe); ONE FIXED set of 2^23
  isotropic rays (uniform_random_rays, seed 1234, direction-sorted) from the box centre with
  length 2*(max_x-min_x) (tests/profile_trace_gadget/profile_trace_gadget.cu:82-109).
One step = one trace_cumulative_sph over the whole ray set, results assembled in ray order.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun, one rank per GPU), STRONG scaling: the same 2^23 rays at every N.  The tree
and the sorted particles are replicated (particles broadcast with NCCL, every rank builds the
same deterministic tree); the rays are dealt to the ranks in 32-aligned tiles of 4096
(round-robin, grace_devel_b200.sharded_trace's split), each rank traces its tiles, the 4 B/ray
results are all-gathered and put back in ray order.  Inside the run rank 0 also traces the
whole set alone and the assembled result must equal it bit for bit.

--impl reference: the reference's own CUDA implementation (GRACE's headers, patched only for
CUDA-12 API removals, oracle/_ref/ref_bench) on the same particles and rays on ONE GPU -- the
reference has no multi-GPU path -- with its host brute force (the reference's own test code)
on a bounded ray sample as `cpu_baseline`.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# torchrun exports OMP_NUM_THREADS=1; the CPU baselines use every core of the box and say how many
HOST_THREADS = os.cpu_count() or 1
os.environ["OMP_NUM_THREADS"] = str(HOST_THREADS)

import numpy as np  # noqa: E402

TILE = 4096
METRIC = "SPH trace Mrays/s (cumulative column density)"
CUM_TOLERANCE = 1e-5          # north_star: column densities within 1e-5 relative


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-particles", type=int, default=24)
    ap.add_argument("--log2-rays", type=int, default=23, help="the whole (fixed) ray set, at every N")
    ap.add_argument("--max-per-leaf", type=int, default=32)
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="0 = auto (~10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-build-timing", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true",
                    help="skip timing the reference's own CUDA build (oracle/_ref/ref_bench[_tuned]) beside this run")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the second block (BASELINE config 5 A: 2^27 particles, 2^24 HEALPix rays)")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def workload_config(args, world):
    return {
        "workload": "profile_trace_gadget: trace_cumulative_sph, 2^%d Gadget-shaped SPH particles, one fixed set of "
                    "2^%d isotropic rays from the box centre (the same set at every GPU count)"
                    % (args.log2_particles, args.log2_rays),
        "particles": 1 << args.log2_particles, "rays": 1 << args.log2_rays,
        "max_per_leaf": args.max_per_leaf, "key_bits": 30, "deltas": "euclidean",
        "parallelism": "rays dealt in 32-aligned %d-ray tiles over %d GPU(s), tree replicated" % (TILE, world),
        "l2": "flushed between timed steps (256 MiB write); inputs (256 MiB spheres + tree + rays) exceed L2",
    }


def kernel_source_sha():
    """Identifies the traversal kernel a stored ncu figure belongs to."""
    h = hashlib.sha1()
    for f in ("trace_packet.cuh", "trace.cu"):
        h.update(open(os.path.join(ROOT, "grace-devel_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


# ----------------------------------------------------------------------------- reference arm
def ref_bench_path(tuned=False):
    return os.path.join(ROOT, "oracle", "_ref", "ref_bench_tuned" if tuned else "ref_bench")


def run_ref_bench(args, steps, warmup, e2e_steps, tuned=False, dump_dir=None, sample=0, timeout=900):
    exe = ref_bench_path(tuned)
    if not os.path.exists(exe):
        return None
    cmd = [exe, str(args.log2_particles), str(args.log2_rays), str(args.max_per_leaf), str(steps), str(warmup),
           str(e2e_steps)]
    if dump_dir:
        cmd += [dump_dir, str(sample)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_bench failed: " + (out.stderr or out.stdout)[-400:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def load_cpu_reference():
    """The reference's own host code (oracle/_ref, built from /root/reference headers) when it is there,
    else the oracle port.  Returns (brute_cumulative_fn, kind, threads)."""
    import oracle
    try:
        from oracle import refcpu
        if refcpu.available():
            return refcpu.brute_cumulative, "reference", refcpu.num_threads()
    except Exception:
        pass
    return oracle.brute_cumulative, "port", oracle.num_threads()


def cpu_sample_size(args, threads):
    # ~15 s of brute force at about 1.5e8 ray-sphere tests per thread-second
    n = 1 << args.log2_particles
    return args.cpu_sample_rays or max(32, min(1 << args.log2_rays, int(15.0 * 1.5e8 * threads / n) // 32 * 32))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n, r = 1 << args.log2_particles, 1 << args.log2_rays
    brute_cum, kind, threads = load_cpu_reference()
    sample_n = cpu_sample_size(args, threads)
    with tempfile.TemporaryDirectory() as d:
        try:
            info = run_ref_bench(args, args.steps, args.warmup, max(1, min(args.steps, 5)), dump_dir=d, sample=sample_n)
        except Exception as e:
            info, err = None, str(e)[:200]
        if info is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_bench did not run: %s"
                              % (err if "err" in dir() else "not built (needs /root/reference at build time)")}))
            return
        h_s = np.fromfile(os.path.join(d, "spheres_sorted.bin"), np.float32).reshape(-1, 4)
        h_r = np.fromfile(os.path.join(d, "rays_sample.bin"), np.float32).reshape(-1, 7)
        got = np.fromfile(os.path.join(d, "cum_sample.bin"), np.float32)
    t0 = time.perf_counter()
    ref = brute_cum(h_r, h_s)
    dt = time.perf_counter() - t0
    rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)))
    ms = info["ms_per_step"]
    val = r / ms / 1e3
    line = {
        "impl": "reference", "reference_kind": "the reference's own CUDA implementation on one GPU (oracle/_ref/ref_bench: "
                                               "GRACE's headers patched only for CUDA-12 API removals, grid cap as shipped)",
        "metric": METRIC, "value": val, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, 1),
        "e2e": {"value": r / info["e2e_ms_per_step"] / 1e3, "unit": "Mrays/s",
                "h2d_bytes_per_step": info["h2d_bytes_per_step"], "d2h_bytes_per_step": info["d2h_bytes_per_step"],
                "what": "particles H2D + tree build + rays H2D + trace + result D2H through the reference's API"},
        "cpu_baseline": {"value": len(h_r) / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
                         "sample": "%d of %d rays (evenly strided), brute force over all %d spheres, %.1f s "
                                   "(tests/tree_traversal/tree_traversal.cu:65-79 pattern)" % (len(h_r), r, n, dt),
                         "max_rel_err_reference_cuda_vs_host": rel,
                         "note": "the host code has no FMA contraction; agreement at the 1e-5 level is expected"},
        "result_fnv1a": info["result_fnv1a"], "n_leaves": info["n_leaves"], "max_blocks": info["max_blocks"],
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- b200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import grace_devel_b200 as gb

    n = 1 << args.log2_particles
    r = 1 << args.log2_rays
    dev = torch.device("cuda", local)
    failures = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- setup (untimed): particles (rank 0 -> NCCL broadcast), tree on every rank ----
    if rank == 0:
        spheres = gb.synth_gadget_spheres(n, 1234)
        h_spheres = spheres.cpu().pin_memory()          # unsorted, as a snapshot reader would deliver them
    else:
        spheres = torch.empty((n, 4), dtype=torch.float32, device=dev)
        h_spheres = None
    if world > 1:
        dist.broadcast(spheres, src=0)

    def build(s):
        tree = gb.Tree(n, args.max_per_leaf)
        gb.build_tree(s, tree)
        return tree

    tree = build(spheres)
    torch.cuda.synchronize()
    lo, hi = gb.min_max_x(spheres)
    c = (hi + lo) / 2.0
    length = 2.0 * (hi - lo)
    # the ray generator is deterministic on a given device type: every rank makes the same set
    rays = torch.empty((r, 7), dtype=torch.float32, device=dev)
    gb.uniform_random_rays(rays, c, c, c, length, 1234)
    local_rays = gb.take_local(rays, rank, world, TILE)
    r_local = local_rays.shape[0]
    out_local = torch.empty(r_local, dtype=torch.float32, device=dev)
    gathered = torch.empty(world * r_local, dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    result = [None]

    def step(src_rays):
        gb.trace_cumulative_sph(src_rays, spheres, tree, out_local)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_local)
            result[0] = gb.scatter_back(gathered, r, world, TILE)
        else:
            result[0] = out_local

    for _ in range(max(args.warmup, 0)):
        step(local_rays)
    barrier()

    # ---- timed region: K steps, device time per step (CUDA events on the launch stream),
    #      L2 flushed between steps; the flush is outside the events ----
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record(stream)
        kern_ev[k][0].record(stream)
        gb.trace_cumulative_sph(local_rays, spheres, tree, out_local)
        kern_ev[k][1].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_local)
            result[0] = gb.scatter_back(gathered, r, world, TILE)
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    kern_ms = [a.elapsed_time(b) for a, b in kern_ev]
    ms_per_step = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / args.steps
    value = r / (ms_per_step * 1e-3) / 1e6
    # a number from a traversal that overflowed its stack or did not terminate is not a number
    derr = gb.device_error()
    if derr != 0:
        raise SystemExit(f"bench: device-side trace error {derr} on rank {rank}")
    final = result[0].clone()

    # ---- the assembled result must equal a single-GPU trace of the whole set, bit for bit ----
    multi_ok = None
    if rank == 0:
        single = torch.empty(r, dtype=torch.float32, device=dev)
        gb.trace_cumulative_sph(rays, spheres, tree, single)
        multi_ok = bool(torch.equal(single.view(torch.int32), final.view(torch.int32)))
        if not multi_ok:
            failures.append("result assembled from %d rank(s) differs from the single-GPU trace" % world)
        result_sha1 = hashlib.sha1(final.cpu().numpy().tobytes()).hexdigest()[:16]
        del single

    # ---- e2e: from HOST buffers -- particles H2D, broadcast, tree build, rays H2D (each rank its share), trace,
    #      gather, result D2H -- every step ----
    e2e_steps = max(1, min(args.steps, 10))
    # every rank holds its share of the rays in pinned host memory and copies it over its own PCIe link (rank 0
    # copying all rays and broadcasting them was 4 ms of every step at any N)
    h_rays_local = gb.take_local(rays, rank, world, TILE).cpu().pin_memory()
    h_out = torch.empty(r, dtype=torch.float32).pin_memory() if rank == 0 else None
    d_rays_local = torch.empty((h_rays_local.shape[0], 7), dtype=torch.float32, device=dev)
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    saved_tree = tree
    for k in range(1 + e2e_steps):
        flush.fill_(k & 0xff)
        barrier()
        if k >= 1:
            e2e_ev[k - 1][0].record(stream)
        if rank == 0:
            spheres.copy_(h_spheres, non_blocking=True)
        if world > 1:
            dist.broadcast(spheres, src=0)
        tree = build(spheres)
        d_rays_local.copy_(h_rays_local, non_blocking=True)
        step(d_rays_local)
        if rank == 0:
            h_out.copy_(result[0], non_blocking=True)
        if k >= 1:
            e2e_ev[k - 1][1].record(stream)
    barrier()
    e2e_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in e2e_ev)) / e2e_steps
    e2e_val = r / (e2e_ms * 1e-3) / 1e6
    if rank == 0 and not np.array_equal(h_out.numpy().view(np.uint32), final.cpu().numpy().view(np.uint32)):
        failures.append("end-to-end result differs from the device-resident one")
    del saved_tree

    # ---- second block: BASELINE config 5 (A): 2^27 particles, 63-bit keys, 2^24 HEALPix rays ----
    config5 = None
    if not args.no_config5:
        try:
            config5 = run_config5(args, gb, dist, world, rank, dev, barrier, max_over_ranks)
        except Exception as e:                          # a second block, never a reason to lose the line
            config5 = {"unavailable": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (trace_packet_kernel<cumulative>, rank 0's launch) ----
    st = gb.trace_stats_sph(local_rays, spheres, tree)
    # B = R*(28 + 4) + 64*node_visits + 16*leaf_visits + 16*prims_staged
    alg_bytes = r_local * 32 + 64 * st["node_visits"] + 16 * st["leaf_visits"] + 16 * st["prims_staged"]
    kern_s = statistics.mean(kern_ms) * 1e-3
    peak, peak_how = measured_peak_gbs()
    achieved = alg_bytes / kern_s / 1e9
    traffic, traffic_note = None, "no ncu capture of this kernel source on file (profiles/trace_traffic.json)"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "trace_traffic.json")))
        if tj.get("kernel_source_sha") == kernel_source_sha() and tj.get("rays_per_launch") == r_local:
            traffic, traffic_note = tj["dram_bytes_per_launch"], "ncu dram__bytes_read+write of this kernel source, " + tj.get("from", "")
        else:
            traffic_note = "profiles/trace_traffic.json is from another kernel source or launch size: not reported"
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "trace_packet_kernel<cumulative,32> (2 launches per call: packets with work stealing, fold)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": traffic_note, "peak_source": peak_how,
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": statistics.mean(kern_ms), "rays_per_launch": r_local,
        "cold_miss_lower_bound_bytes": 32 * r_local + 16 * n + 64 * (tree.n_leaves - 1) + 16 * tree.n_leaves,
        "ray_sphere_tests_per_s": 32.0 * st["prims_staged"] / kern_s,
        "hits_per_ray": st["hits"] / r_local,
        "note": "traversal is L2/issue-bound: algorithmic bytes count every node/leaf fetch of the "
                "reference packet algorithm, most of which hit L2",
    }

    # ---- tree build (secondary metric of BASELINE.json: LBVH build Mparticles/s) ----
    build_info = None
    if not args.no_build_timing:
        work = torch.empty_like(spheres)
        times = []
        d_unsorted = h_spheres.to(dev)
        for k in range(4):
            work.copy_(d_unsorted)
            t_tree = gb.Tree(n, args.max_per_leaf)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(k)
            a.record(stream)
            gb.build_tree(work, t_tree)
            b.record(stream)
            torch.cuda.synchronize()
            if k:
                times.append(a.elapsed_time(b))
            del t_tree
        bms = statistics.mean(times)
        L = tree.n_leaves
        bbytes = n * (116 + 88.0 * L / n)
        build_info = {"value": n / (bms * 1e-3) / 1e6, "unit": "Mparticles/s", "ms": bms,
                      "n_leaves": L, "algorithmic_bytes": bbytes,
                      "hbm_frac": bbytes / (bms * 1e-3) / 1e9 / peak,
                      "stages": "bounds + 30-bit keys + onesweep sort + Euclidean deltas + leaves + nodes"}
        del work, d_unsorted

    # ---- CPU baseline: brute force on a bounded ray sample, also a full-size parity check ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle
        brute_cum, kind, threads = load_cpu_reference()
        sample_n = cpu_sample_size(args, threads)
        idx = torch.arange(sample_n, device=dev) * (r // sample_n)
        h_s = spheres.cpu().numpy()
        h_r = rays[idx].cpu().numpy()
        t0 = time.perf_counter()
        ref = brute_cum(h_r, h_s)
        dt = time.perf_counter() - t0
        got = final[idx].cpu().numpy()
        # the parity figure is against the oracle, which restates the DEVICE arithmetic (FMA contraction):
        # it must agree bit for bit, also at full size.  The reference's HOST code has no FMA contraction and
        # is quoted for information.
        n_exact = min(sample_n, 512)
        exact = oracle.brute_cumulative(h_r[:n_exact], h_s)
        rel_oracle = float(np.max(np.abs(got[:n_exact] - exact) / np.maximum(np.abs(exact), 1e-30)))
        bit_exact = bool(np.array_equal(got[:n_exact].view(np.uint32), exact.view(np.uint32)))
        if rel_oracle > CUM_TOLERANCE:
            failures.append("column densities differ from the oracle by %.3g relative (tolerance %.0e)" % (rel_oracle, CUM_TOLERANCE))
        cpu = {"value": sample_n / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
               "sample": "%d of %d rays (evenly strided), brute force over all %d spheres, %.1f s"
                         % (sample_n, r, n, dt),
               "parity_max_rel_err_vs_oracle": rel_oracle, "parity_tolerance": CUM_TOLERANCE,
               "parity_bit_exact_vs_oracle": bit_exact, "parity_rays_checked": n_exact,
               "max_rel_err_vs_reference_host_code": float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30))),
               "note": "the reference's host code is compiled without FMA contraction, the device code with: "
                       "their agreement (last figure) is informational, the contract is checked against the oracle"}

    # ---- the reference's own CUDA build on the same box and inputs (reported baselines) ----
    ref_cuda = None
    if not args.no_reference_cuda and world == 1:
        ref_cuda = {}
        for name, tuned in (("as_shipped", False), ("tuned", True)):
            try:
                with tempfile.TemporaryDirectory() as d:
                    info = run_ref_bench(args, 3, 1, 2, tuned=tuned, dump_dir=d, sample=4096)
                    same = None
                    if info is not None:
                        ref_sample = np.fromfile(os.path.join(d, "cum_sample.bin"), np.float32)
                        ours = final[torch.arange(4096, device=dev) * (r // 4096)].cpu().numpy()
                        same = bool(np.array_equal(ref_sample.view(np.uint32), ours.view(np.uint32)))
                        if not same:
                            failures.append("column densities differ from the reference CUDA build's (%s)" % name)
                if info is None:
                    ref_cuda[name] = {"unavailable": "oracle/_ref/%s not built" % os.path.basename(ref_bench_path(tuned))}
                    continue
                ref_cuda[name] = {"value": r / info["ms_per_step"] / 1e3, "unit": "Mrays/s", "ms_per_step": info["ms_per_step"],
                                  "e2e_value": r / info["e2e_ms_per_step"] / 1e3, "e2e_ms_per_step": info["e2e_ms_per_step"],
                                  "max_blocks": info["max_blocks"], "sample_of_4096_rays_bit_identical": same,
                                  "speedup_of_this_repo": info["ms_per_step"] / ms_per_step,
                                  "e2e_speedup_of_this_repo": info["e2e_ms_per_step"] / e2e_ms}
            except Exception as e:      # a baseline, never a reason to lose the bench line
                ref_cuda[name] = {"unavailable": str(e)[:200]}
        ref_cuda["what"] = ("GRACE's headers (patched only for CUDA-12 API removals, oracle/patch_ref.py) through its public API on "
                            "the same particles and rays: grid cap as shipped (kernel_config.h:11 MAX_BLOCKS = 112) and lifted to "
                            "148 x 8 blocks ('tuned'); CUDA events, 3 steps")

    # ---- BASELINE configs[2] literally (2^20 rays, the round-1 bench workload), next to the same two reference builds ----
    small = None
    if world == 1 and args.log2_rays != 20:
        r20 = 1 << 20
        rays20 = torch.empty((r20, 7), dtype=torch.float32, device=dev)
        gb.uniform_random_rays(rays20, c, c, c, length, 1234)
        out20 = torch.empty(r20, dtype=torch.float32, device=dev)
        ts = []
        for k in range(8):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            gb.trace_cumulative_sph(rays20, spheres, tree, out20)
            b.record(stream)
            torch.cuda.synchronize()
            if k >= 3:
                ts.append(a.elapsed_time(b))
        ms20 = statistics.mean(ts)
        small = {"workload": "2^24 particles, 2^20 isotropic rays (BASELINE configs[2] as written; BENCH_r01's workload)",
                 "ms_per_step": ms20, "value": r20 / ms20 / 1e3, "unit": "Mrays/s"}
        if not args.no_reference_cuda:
            class A20:      # the same reference programs on the smaller ray set
                log2_particles, log2_rays, max_per_leaf = args.log2_particles, 20, args.max_per_leaf
            for name, tuned in (("reference_cuda_as_shipped", False), ("reference_cuda_tuned", True)):
                try:
                    info = run_ref_bench(A20, 3, 1, 1, tuned=tuned)
                    if info is not None:
                        small[name] = {"ms_per_step": info["ms_per_step"], "value": r20 / info["ms_per_step"] / 1e3,
                                       "speedup_of_this_repo": info["ms_per_step"] / ms20}
                except Exception as e:
                    small[name] = {"unavailable": str(e)[:200]}
        del rays20, out20

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Mrays/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": int(n * 16 + r * 28), "d2h_bytes_per_step": int(r * 4),
                "what": "particles H2D + NCCL broadcast + tree build on every rank + H2D of each rank's share of the rays + trace + "
                        "all-gather + reassembly + result D2H, every step"},
        # trace_cumulative_sph = 2 launches of trace_packet_kernel (packets with work stealing, fold) per rank
        "gpu_launches": args.steps * world * 2,
        "parity": {"assembled_equals_single_gpu_bitwise": multi_ok, "result_sha1_16": result_sha1,
                   "failures": failures},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reference_cuda": ref_cuda,
        "build": build_info,
        "config3_2p20_rays": small,
        "config5": config5,
        "wall_s_timed_region": t_wall,
        "n_leaves": tree.n_leaves,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if failures:
        raise SystemExit("bench: parity check failed: " + "; ".join(failures))


def run_config5(args, gb, dist, world, rank, dev, barrier, max_over_ranks):
    """BASELINE config 5 (A): 2^27 Gadget-shaped particles, 63-bit keys, HEALPix NESTED pixels [0, 2^24) of
    nside 2048 from the box centre; tree replicated, rays sharded.  End to end per step: tree build on rank 0,
    NCCL broadcast of the finished tree, trace of the rank's tiles, all-gather + reassembly."""
    import torch
    n5, r5 = 1 << 27, 1 << 24
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        return {"unavailable": "needs ~40 GiB free device memory, %.1f GiB free" % (free / 2 ** 30)}
    stream = torch.cuda.current_stream()
    src = gb.synth_gadget_spheres(n5, 1234) if rank == 0 else None
    s5 = torch.empty((n5, 4), dtype=torch.float32, device=dev)
    box = torch.zeros(2, dtype=torch.float64, device=dev)
    if rank == 0:
        lo, hi = gb.min_max_x(src)
        box[0], box[1] = lo, hi
    if world > 1:
        dist.broadcast(box, src=0)
    lo, hi = float(box[0]), float(box[1])
    c = (hi + lo) / 2.0
    rays5 = gb.healpix_rays(None, 2048, 0, r5, c, c, c, 2.0 * (hi - lo))
    local5 = gb.take_local(rays5, rank, world, TILE)
    out5 = torch.empty(local5.shape[0], dtype=torch.float32, device=dev)
    gath5 = torch.empty(world * local5.shape[0], dtype=torch.float32, device=dev) if world > 1 else None
    res, t_b, t_t, t_all = None, [], [], []
    for k in range(3):
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(stream)
        # The tree is built once, on rank 0, and the finished tree is broadcast (sorted particles 2 GiB +
        # nodes 0.43 GB + leaves 0.1 GB over NVLink): with every rank building its own copy the build was the
        # Amdahl term of this configuration -- and ran 2x slower per rank than alone.
        tree5 = gb.Tree(2, args.max_per_leaf) if rank else gb.Tree(n5, args.max_per_leaf)
        # build_tree (tests/helper/tree.cuh:15-43) in its three steps: the sorted particles are final after the first, so
        # their broadcast (2 GiB, the bulk of the tree) runs on NCCL's stream while rank 0 computes deltas, leaves and nodes
        sent = None
        if rank == 0:
            s5.copy_(src)
            gb.morton_keys63_sort_sph(s5)
        if world > 1:
            sent = dist.broadcast(s5, src=0, async_op=True)
        if rank == 0:
            deltas5 = torch.empty(n5 + 1, dtype=torch.float32, device=dev)
            gb.euclidean_deltas_sph(s5, deltas5)
            gb.ALBVH_sph(s5, deltas5, tree5)
            del deltas5
        e[1].record(stream)
        if world > 1:
            nl = torch.tensor([tree5.n_leaves if rank == 0 else 0], dtype=torch.int64, device=dev)
            dist.broadcast(nl, src=0)
            L = int(nl.item())
            if rank:
                tree5.nodes = torch.empty((L - 1, 16), dtype=torch.int32, device=dev)
                tree5.leaves = torch.empty((L, 4), dtype=torch.int32, device=dev)
            for t in (tree5.nodes, tree5.leaves, tree5.root_index_ptr):
                dist.broadcast(t, src=0)
            sent.wait()
        e[2].record(stream)
        gb.trace_cumulative_sph(local5, s5, tree5, out5)
        if world > 1:
            dist.all_gather_into_tensor(gath5, out5)
            res = gb.scatter_back(gath5, r5, world, TILE)
        else:
            res = out5
        e[3].record(stream)
        barrier()
        if k:
            t_b.append(max_over_ranks(e[0].elapsed_time(e[2])))
            t_t.append(max_over_ranks(e[2].elapsed_time(e[3])))
            t_all.append(max_over_ranks(e[0].elapsed_time(e[3])))
    if gb.device_error() != 0:
        raise RuntimeError("device-side trace error in config 5")
    sha = hashlib.sha1(res.cpu().numpy().tobytes()).hexdigest()[:16] if rank == 0 else None
    n_leaves = tree5.n_leaves
    del tree5, s5, src, rays5
    torch.cuda.empty_cache()
    tt, ta = statistics.mean(t_t), statistics.mean(t_all)
    return {"workload": "one_to_many_rays (A): 2^27 particles, 63-bit keys, 2^24 HEALPix NESTED rays (nside 2048, pixels [0, 2^24))",
            "n_gpus": world, "build_and_tree_broadcast_ms": statistics.mean(t_b), "trace_gather_ms": tt,
            "end_to_end_ms": ta, "mrays_s_trace_gather": r5 / tt / 1e3, "mrays_s_end_to_end": r5 / ta / 1e3,
            "end_to_end": "tree build on rank 0 + NCCL broadcast of the finished tree (sorted particles -- sent while deltas, leaves and "
                          "nodes are computed --, nodes, leaves, root) + trace of the rank's tiles + all-gather + reassembly",
            "n_leaves": n_leaves, "result_sha1_16": sha}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
