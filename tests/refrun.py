"""Run the reference's own CUDA implementation (oracle/_ref/ref_driver) on given inputs and
load what it dumps.  Test infrastructure."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def available():
    return os.path.exists(DRIVER)


def run(spheres, rays, max_per_leaf=32, key_bits=30, iters=0, lists=True, timeout=600):
    """rays: [R,7] float32 array, or a 'gen:N:seed:ox:oy:oz:len' string (the reference's
    uniform_random_rays).  Returns (dict of arrays, timing dict)."""
    with tempfile.TemporaryDirectory() as d:
        sp = os.path.join(d, "spheres.bin")
        np.ascontiguousarray(spheres, np.float32).tofile(sp)
        if isinstance(rays, str):
            rarg = rays
        else:
            rarg = os.path.join(d, "rays_in.bin")
            np.ascontiguousarray(rays, np.float32).tofile(rarg)
        cmd = [DRIVER, sp, rarg, d, str(max_per_leaf), str(key_bits), str(iters)] + (["lists"] if lists else [])
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if out.returncode != 0:
            raise RuntimeError("ref_driver failed: " + out.stderr[-2000:] + out.stdout[-500:])
        info = json.loads(out.stdout.strip().splitlines()[-1])

        def ld(name, dt, shape=None):
            a = np.fromfile(os.path.join(d, name), dtype=dt)
            return a.reshape(shape) if shape else a
        res = {
            "spheres_sorted": ld("spheres_sorted.bin", np.float32, (-1, 4)),
            "deltas": ld("deltas.bin", np.float32),
            "leaves": ld("leaves.bin", np.int32, (-1, 4)),
            "nodes": ld("nodes.bin", np.int32, (-1, 16)),
            "root": int(ld("root.bin", np.int32)[0]),
            "rays": ld("rays.bin", np.float32, (-1, 7)),
            "hitcounts": ld("hitcounts.bin", np.int32),
            "cumulative": ld("cumulative.bin", np.float32),
        }
        if lists:
            res.update(offsets=ld("offsets.bin", np.int32), hit_idx=ld("hit_idx.bin", np.int32),
                       hit_integral=ld("hit_integral.bin", np.float32), hit_dist=ld("hit_dist.bin", np.float32))
        return res, info


GEN_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_gen_driver")
GEN_NAMES = ("octant", "o2m_nosort", "o2m_dirsort", "o2m_endsort_aabb", "plane_parallel", "ortho", "pinhole")
# the parameters ref_gen_driver.cu hard-codes (everything but sizes and seed)
GEN_PARAMS = dict(octant_origin=(0.5, 0.25, 0.125), octant_length=2.0, octant="MPM",
                  o2m_origin=(0.1, 0.2, 0.3), o2m_aabb=((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
                  pp_base=(-0.1, -0.2, 1.5), pp_w=(1.3, 0.0, 0.0), pp_h=(0.0, 1.1, 0.0), pp_length=3.0,
                  cam_pos=(0.5, 0.4, 2.0), look_at=(0.45, 0.5, 0.5), view_up=(0.0, 1.0, 0.1),
                  ortho_extent=1.25, pinhole_fovy=0.9, cam_length=4.0)


def gens_available():
    return os.path.exists(GEN_DRIVER)


def run_gens(points, n_random, seed, res_x, res_y, timeout=300):
    """Every ray generator of the reference on this device.  points: [P,3] float32 end points."""
    with tempfile.TemporaryDirectory() as d:
        pp = os.path.join(d, "points.bin")
        np.ascontiguousarray(points, np.float32).tofile(pp)
        out = subprocess.run([GEN_DRIVER, d, pp, str(n_random), str(seed), str(res_x), str(res_y)],
                             capture_output=True, text=True, timeout=timeout)
        if out.returncode != 0:
            raise RuntimeError("ref_gen_driver failed: " + out.stderr[-2000:] + out.stdout[-500:])
        return {k: np.fromfile(os.path.join(d, k + ".bin"), dtype=np.float32).reshape(-1, 7) for k in GEN_NAMES}
