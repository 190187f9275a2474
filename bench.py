#!/usr/bin/env python
"""bench.py -- GRACE hot path on B200: SPH cumulative column-density trace (Mrays/s).

Workload (BASELINE.json configs[2], "profile_trace_gadget", the configuration the metric
"SPH trace Mrays/s (2^24 particles ...)" is quoted on; it fits one GPU):
  2^24 synthetic Gadget-shaped SPH particles (float4 x,y,z,h), ALBVH with max_per_leaf=32,
  30-bit keys, Euclidean deltas (tests/helper/tree.cuh:15-27 recipe); 2^20 isotropic rays
  (uniform_random_rays, seed 1234, direction-sorted) from the box centre with length
  2*(max_x-min_x) (tests/profile_trace_gadget/profile_trace_gadget.cu:82-109).
One step = one trace_cumulative_sph over all rays of the rank.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun, one rank per GPU): particles are broadcast with NCCL, every rank builds
the same (deterministic) tree and traces ITS OWN 2^20 rays (seed 1234 + rank): weak
scaling, no collective on the data path except the gather of 4 B/ray results to rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-particles", type=int, default=24)
    ap.add_argument("--log2-rays", type=int, default=20)
    ap.add_argument("--max-per-leaf", type=int, default=32)
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="0 = auto (~10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-build-timing", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true",
                    help="skip timing the reference's own CUDA build (oracle/_ref/ref_driver) beside this run")
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


# ----------------------------------------------------------------------------- reference arm
def load_cpu_reference():
    """The reference's own CPU code (oracle/_ref, built from /root/reference headers) when
    it is there, else the oracle port.  Returns (brute_cumulative_fn, kind, threads)."""
    import oracle
    try:
        from oracle import refcpu
        if refcpu.available():
            return refcpu.brute_cumulative, refcpu.brute_hitcounts, "reference", refcpu.num_threads()
    except Exception:
        pass
    return oracle.brute_cumulative, oracle.brute_hitcounts, "port", oracle.num_threads()


def host_workload(args, seed_rays=1234):
    """Particles + rays on the host for the CPU arms.  Uses the CUDA generators when a GPU
    is present (same bits as the b200 arm), numpy otherwise."""
    n = 1 << args.log2_particles
    r = 1 << args.log2_rays
    try:
        import torch
        if torch.cuda.is_available():
            import grace_devel_b200 as gb
            s = gb.synth_gadget_spheres(n, 1234)
            lo, hi = gb.min_max_x(s)
            c = (hi + lo) / 2.0
            rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
            gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), seed_rays)
            return s.cpu().numpy(), rays.cpu().numpy(), "cuda generators"
    except Exception:
        pass
    from util import clustered_spheres, isotropic_rays
    s = clustered_spheres(n, seed=1234, n_halos=64)
    lo, hi = float(s[:, 0].min()), float(s[:, 0].max())
    c = (hi + lo) / 2.0
    rays = isotropic_rays(r, origin=(c, c, c), length=2 * (hi - lo), seed=seed_rays)
    return s, rays, "numpy generators (no GPU)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    brute_cum, _, kind, threads = load_cpu_reference()
    s, rays, how = host_workload(args)
    n, r = len(s), len(rays)
    # bounded sample per step: ~2 s of work (about 1.5e8 ray-sphere tests per thread-second)
    per_step = max(32, int(2.0 * 1.5e8 * threads / n) // 32 * 32)
    per_step = min(per_step, r)
    stride = max(1, r // per_step)

    def sample(k):
        idx = (np.arange(per_step) * stride + k) % r
        return np.ascontiguousarray(rays[idx])

    for k in range(args.warmup):
        brute_cum(sample(k), s)
    t0 = time.perf_counter()
    for k in range(args.steps):
        brute_cum(sample(args.warmup + k), s)
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = per_step / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference",
        "metric": "SPH trace Mrays/s (cumulative column density)", "value": val, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (%s)" % how,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": kind,
                         "sample": "%d of %d rays per step, brute force over all %d spheres "
                                   "(tests/tree_traversal/tree_traversal.cu:65-79 pattern)" % (per_step, r, n)},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {
        "workload": "profile_trace_gadget: trace_cumulative_sph, 2^%d Gadget-shaped SPH particles, "
                    "2^%d isotropic rays per GPU from the box centre" % (args.log2_particles, args.log2_rays),
        "particles": 1 << args.log2_particles, "rays_per_gpu": 1 << args.log2_rays,
        "max_per_leaf": args.max_per_leaf, "key_bits": 30, "deltas": "euclidean",
        "parallelism": "rays sharded over %d GPU(s), tree replicated" % world,
        "l2": "flushed between timed steps (256 MiB write); inputs (256 MiB spheres + tree) exceed L2",
    }


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

    # ---- setup (untimed): particles (rank 0 -> NCCL broadcast), tree on every rank ----
    if rank == 0:
        spheres = gb.synth_gadget_spheres(n, 1234)
    else:
        spheres = torch.empty((n, 4), dtype=torch.float32, device=dev)
    if world > 1:
        dist.broadcast(spheres, src=0)
    unsorted = spheres.clone() if not args.no_build_timing and rank == 0 else None

    def build(s):
        tree = gb.Tree(n, args.max_per_leaf)
        gb.build_tree(s, tree)
        return tree

    tree = build(spheres)
    torch.cuda.synchronize()
    lo, hi = gb.min_max_x(spheres)
    c = (hi + lo) / 2.0
    length = 2.0 * (hi - lo)
    rays = torch.empty((r, 7), dtype=torch.float32, device=dev)
    gb.uniform_random_rays(rays, c, c, c, length, 1234 + rank)
    out = torch.empty(r, dtype=torch.float32, device=dev)
    gathered = torch.empty(world * r, dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h_rays = rays.cpu().pin_memory()
    h_out = torch.empty(r, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream()

    def step():
        gb.trace_cumulative_sph(rays, spheres, tree, out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()

    # ---- timed region: K steps, device time per step (CUDA events on the launch stream),
    #      L2 flushed between steps; the flush is outside the events ----
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record(stream)
        kern_ev[k][0].record(stream)
        gb.trace_cumulative_sph(rays, spheres, tree, out)
        kern_ev[k][1].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kern_ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * r / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the reference-facing call with HOST buffers (H2D rays, trace, D2H result) ----
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
    d_rays2 = torch.empty_like(rays)
    for k in range(2 + args.steps):
        flush.fill_(k & 0xff)
        if k >= 2:
            e2e_ev[k - 2][0].record(stream)
        d_rays2.copy_(h_rays, non_blocking=True)
        gb.trace_cumulative_sph(d_rays2, spheres, tree, out)
        h_out.copy_(out, non_blocking=True)
        if k >= 2:
            e2e_ev[k - 2][1].record(stream)
    barrier()
    e2e_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e_ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * r / (float(e2e_ms.item()) / args.steps * 1e-3) / 1e6
    # a number from a traversal that overflowed its stack or did not terminate is not a number
    derr = gb.device_error()
    if derr != 0:
        raise SystemExit(f"bench: device-side trace error {derr} on rank {rank}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (trace_kernel<cumulative>) ----
    st = gb.trace_stats_sph(rays, spheres, tree)
    # B = R*(28 + 4) + 64*node_visits + 16*leaf_visits + 16*prims_staged
    alg_bytes = r * 32 + 64 * st["node_visits"] + 16 * st["leaf_visits"] + 16 * st["prims_staged"]
    kern_s = statistics.mean(kern_ms) * 1e-3
    peak, peak_how = measured_peak_gbs()
    achieved = alg_bytes / kern_s / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "trace_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "trace_packet_kernel<cumulative,32> (4 launches per call: packets + 3 load-balancing rounds)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_how,
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": statistics.mean(kern_ms),
        "cold_miss_lower_bound_bytes": 32 * r + 16 * n + 64 * (tree.n_leaves - 1) + 16 * tree.n_leaves,
        "ray_sphere_tests_per_s": 32.0 * st["prims_staged"] / kern_s,
        "hits_per_ray": st["hits"] / r,
        "note": "traversal is L2/issue-bound: algorithmic bytes count every node/leaf fetch of the "
                "reference packet algorithm, most of which hit L2",
    }

    # ---- tree build (secondary metric of BASELINE.json: LBVH build Mparticles/s) ----
    build_info = None
    if unsorted is not None:
        work = torch.empty_like(unsorted)
        times = []
        for k in range(4):
            work.copy_(unsorted)
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
        del work

    # ---- CPU baseline: brute force on a bounded ray sample, also a full-size parity check ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        brute_cum, brute_cnt, kind, threads = load_cpu_reference()
        sample_n = args.cpu_sample_rays or max(32, min(r, int(15.0 * 1.5e8 * threads / n) // 32 * 32))
        idx = torch.arange(sample_n, device=dev) * (r // sample_n)
        h_s = spheres.cpu().numpy()
        h_r = rays[idx].cpu().numpy()
        t0 = time.perf_counter()
        ref = brute_cum(h_r, h_s)
        dt = time.perf_counter() - t0
        got = out[idx].cpu().numpy()
        rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)))
        # the reference's HOST code has no FMA contraction (1e-5 agreement); the oracle port
        # restates the DEVICE arithmetic and must agree bit for bit, also at full size
        import oracle
        n_exact = min(sample_n, 512)
        exact = oracle.brute_cumulative(h_r[:n_exact], h_s)
        cpu = {"value": sample_n / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
               "sample": "%d of %d rays (evenly strided), brute force over all %d spheres, %.1f s"
                         % (sample_n, r, n, dt),
               "parity_max_rel_err_vs_gpu": rel,
               "parity_bit_exact_vs_oracle": bool(np.array_equal(got[:n_exact].view(np.uint32), exact.view(np.uint32))),
               "parity_rays_checked_bit_exact": n_exact}

    # ---- the reference's own CUDA build on the same box and inputs (a reported baseline) ----
    ref_cuda = None
    if not args.no_reference_cuda and world == 1:
        try:
            import refrun
            if refrun.available():
                h_in = gb.synth_gadget_spheres(n, 1234).cpu().numpy()
                _, info = refrun.run(h_in, "gen:%d:1234:%.9g:%.9g:%.9g:%.9g" % (r, c, c, c, length),
                                     args.max_per_leaf, 30, iters=5, lists=False, timeout=600)
                best = info.get("min_ms") or {"cumulative": info["ms_cumulative"], "hitcounts": info["ms_hitcounts"],
                                              "keys_sort": info["ms_keys_sort"], "deltas": info["ms_deltas"],
                                              "albvh": info["ms_albvh"]}
                ref_ms = best["cumulative"]
                ref_build = best["keys_sort"] + best["deltas"] + best["albvh"]
                ref_cuda = {"value": r / ref_ms / 1e3, "unit": "Mrays/s", "ms_cumulative": ref_ms,
                            "ms_hitcounts": best["hitcounts"], "build_ms": ref_build,
                            "build_mparticles_s": n / ref_build / 1e3, "timing": "best of 5 iterations per stage",
                            "what": "GRACE's headers (patched only for CUDA-12 API removals, oracle/patch_ref.py) "
                                    "called through its public API on the same particles and rays, CUDA events"}
        except Exception as e:      # a baseline, never a reason to lose the bench line
            ref_cuda = {"unavailable": str(e)[:200]}

    line = {
        "metric": "SPH trace Mrays/s (cumulative column density)", "value": value, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": int(h_rays.numel() * 4),
                "d2h_bytes_per_step": int(h_out.numel() * 4)},
        # trace_cumulative_sph = 4 launches of trace_packet_kernel (packets + 3 load-balancing rounds)
        "gpu_launches": args.steps * world * 4,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reference_cuda": ref_cuda,
        "build": build_info,
        "wall_s_timed_region": t_wall,
        "n_leaves": tree.n_leaves,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
