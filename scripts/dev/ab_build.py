"""Dev probe: time the tree build stages at 2^24 (not product, not a test)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import grace_devel_b200 as gb
lg_n = int(os.environ.get("AB_LOG2_N", "24")); n = 1 << lg_n
bits = int(os.environ.get("AB_BITS", "30"))
s0 = gb.synth_gadget_spheres(n, 1234)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
res = {"lib": os.environ.get("GRACE_B200_LIB", "default").split("_")[-1], "n": n, "bits": bits}
acc = {"sort": 0.0, "deltas": 0.0, "albvh": 0.0, "total": 0.0}
reps = int(os.environ.get("AB_REPS", "5"))
for k in range(reps + 1):
    s = s0.clone(); tree = gb.Tree(n, 32); deltas = torch.empty(n + 1, dtype=torch.float32, device="cuda")
    flush.fill_(k); e = [ev() for _ in range(4)]
    e[0].record()
    (gb.morton_keys30_sort_sph if bits == 30 else gb.morton_keys63_sort_sph)(s)
    e[1].record()
    gb.euclidean_deltas_sph(s, deltas)
    e[2].record()
    gb.ALBVH_sph(s, deltas, tree)
    e[3].record(); torch.cuda.synchronize()
    if k:
        acc["sort"] += e[0].elapsed_time(e[1]); acc["deltas"] += e[1].elapsed_time(e[2])
        acc["albvh"] += e[2].elapsed_time(e[3]); acc["total"] += e[0].elapsed_time(e[3])
for k_ in acc: res[k_ + "_ms"] = acc[k_] / reps
res["n_leaves"] = tree.n_leaves
import hashlib
res["nodes_sha"] = hashlib.sha1(tree.nodes.cpu().numpy().tobytes()).hexdigest()[:12]
res["mparticles_s"] = n / res["total_ms"] / 1e3
print(json.dumps(res))
