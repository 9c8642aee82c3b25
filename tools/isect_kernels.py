"""Runs the tile-intersection pipeline twice on the bench workload (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import gsplat_b200
from gsplat_b200 import ops
from tests import scene

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
sc = scene.make_scene(scene_grid=3, sh_degree=3)
dev = torch.device("cuda:0")
P = {k: torch.from_numpy(sc[k]).to(dev) for k in ("means", "quats", "scales", "opacities", "sh")}
P["scales"] = P["scales"] * scale
W, H = 1920, 1080
K = torch.from_numpy(scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)[:1]).to(dev)
vm = torch.from_numpy(sc["viewmats"][:1]).to(dev)
radii, m2, dep, con, col, _ = ops.fused_project_sh(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], vm, K, W, H, 3)
op = P["opacities"][None].contiguous()
for _ in range(2):
    a = ops.isect_tiles(m2, radii, dep, 16, 120, 68, conics=con, opacities=op)
    off = ops.isect_offset_encode(a[1], 1, 120, 68)
torch.cuda.synchronize()
u = ops.isect_tiles(m2, radii, dep, 16, 120, 68, conics=con, opacities=op, sort=False)
ids, perm = torch.sort(u[1], stable=True)
print("n_isects", a[1].numel(), "equal", torch.equal(a[1], ids), torch.equal(a[2], u[2][perm]))
