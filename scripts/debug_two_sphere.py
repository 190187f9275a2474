import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import grace_devel_b200 as gb
from util import ortho_rays_z
mode = sys.argv[1]; nside = int(sys.argv[2])
gb.set_trace_mode(mode)
radius = 0.2
s = np.array([[-0.5, -0.5, -0.5, radius], [0.5, 0.5, 0.5, radius]], np.float32)
d_s = torch.from_numpy(s).cuda()
tree = gb.Tree(2, 1)
gb.build_tree(d_s, tree, -np.ones(3, np.float32), np.ones(3, np.float32))
torch.cuda.synchronize()
print("tree", tree.nodes.cpu().numpy().tolist(), tree.leaves.cpu().numpy().tolist(), int(tree.root_index_ptr.item()), flush=True)
span = 2.0 + 2 * radius
rays = ortho_rays_z(nside, -1.0 - radius, 1.0 + radius)
rays[:, 5] = 1.0 + radius; rays[:, 6] = 2 * span
d_r = torch.from_numpy(rays).cuda()
out = torch.zeros(len(rays), dtype=torch.float32, device="cuda")
cnt = torch.zeros(len(rays), dtype=torch.int32, device="cuda")
gb.trace_hitcounts_sph(d_r, d_s, tree, cnt); torch.cuda.synchronize()
print("counts ok", int(cnt.sum()), "err", gb.device_error(), flush=True)
gb.trace_cumulative_sph(d_r, d_s, tree, out); torch.cuda.synchronize()
print("cum ok", float(out.double().sum()) * (span / nside) ** 2 / 2, "err", gb.device_error(), flush=True)
