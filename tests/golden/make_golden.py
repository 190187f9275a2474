#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by running the REFERENCE'S OWN CUDA
implementation (oracle/_ref/ref_driver, see oracle/build_ref.sh) on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

Each file holds the inputs (spheres, rays) and every output of the reference's hot path for
them; tests/test_oracle_golden.py replays the inputs through the CPU oracle (no GPU needed)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refrun  # noqa: E402
from util import clustered_spheres, uniform_spheres, isotropic_rays, ortho_rays_z  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else HERE
os.makedirs(out, exist_ok=True)

cases = {
    "clustered_mpl32_k30": (clustered_spheres(3000, seed=101, n_halos=4), isotropic_rays(256, seed=5), 32, 30),
    "clustered_mpl8_k63": (clustered_spheres(2500, seed=102, n_halos=3), isotropic_rays(128, seed=6), 8, 63),
    "uniform_mpl1_k30": (uniform_spheres(1500, seed=103, rmax=0.05), isotropic_rays(128, seed=7), 1, 30),
    "uniform_mpl4_ortho": (uniform_spheres(2000, seed=104, rmax=0.04), ortho_rays_z(16), 4, 30),
    "duplicates_mpl32_k30": (np.repeat(uniform_spheres(300, seed=105, rmax=0.08), 8, axis=0),
                             isotropic_rays(64, seed=8), 32, 30),
}
for name, (s, rays, mpl, bits) in cases.items():
    res, info = refrun.run(s, rays, mpl, bits, iters=0, lists=True)
    np.savez_compressed(os.path.join(out, name + ".npz"), spheres=s, rays_in=rays, max_per_leaf=mpl,
                        key_bits=bits, **res)
    print(name, "leaves", len(res["leaves"]), "hits", len(res["hit_idx"]))
# the reference's random ray generator on this device (148 SMs)
res, info = refrun.run(uniform_spheres(64, seed=1), "gen:4096:1234:0.5:0.5:0.5:2.0", 32, 30, iters=0, lists=False)
np.savez_compressed(os.path.join(out, "uniform_random_rays_4096_seed1234_b200.npz"), rays=res["rays"])
# every ray generator of the reference (deterministic ones are device-independent)
pts = np.random.default_rng(21).random((1500, 3), dtype=np.float32)
gens = refrun.run_gens(pts, 2048, 1234, 48, 32)
np.savez_compressed(os.path.join(out, "ray_generators_ref.npz"), points=pts, n_random=2048, seed=1234,
                    res_x=48, res_y=32, **gens)
print("done")
