#!/usr/bin/env python
"""Lane efficiency of the leaf tests (needs a -DPK_PROF_LANES build, GRACE_B200_LIB=...)."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grace_devel_b200 as gb
n = 1 << 24; r = 1 << int(os.environ.get("AB_LOG2_R", 20))
s = gb.synth_gadget_spheres(n, 1234); tree = gb.Tree(n, 32); gb.build_tree(s, tree)
lo, hi = gb.min_max_x(s); c = (lo + hi) / 2
rays = torch.empty((r, 7), dtype=torch.float32, device="cuda")
gb.uniform_random_rays(rays, c, c, c, 2 * (hi - lo), seed=1234)
pp = gb.trace_packet_profile_sph(rays, s, tree)
per = pp.pop("per_packet")
dense_useful, dense_kept, sparse_pairs = per[:, 1].sum(), per[:, 2].sum(), per[:, 3].sum()
print(json.dumps(dict(pp, dense_lane_tests=int(dense_kept * 32), dense_useful=int(dense_useful),
      dense_lane_eff=float(dense_useful / (dense_kept * 32)), sparse_pair_tests=int(sparse_pairs),
      dense_share_of_kept=float(dense_kept / pp["prims_kept"]))))
