import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, hashlib
import grace_devel_b200 as gb
n = 1 << 24
s0 = gb.synth_gadget_spheres(n, 1234)
ts = []
for k in range(6):
    s = s0.clone()
    tree = gb.Tree(n, 32)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gb.build_tree(s, tree); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
h = hashlib.sha1()
L = tree.n_leaves if hasattr(tree, "n_leaves") else None
for t in (tree.nodes, tree.leaves):
    h.update(t.cpu().numpy().tobytes())
print(json.dumps({"build_ms": [round(t, 3) for t in ts], "tree_sha": h.hexdigest()[:12]}))
