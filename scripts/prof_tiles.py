"""Minimal driver for ncu: prepare config 4 (or argv mesh/grid) and voxelize it a few times through the prepared-mesh path."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuda_voxelizer_b200 import _lib
if os.environ.get("VOXB200_SO"):
    _lib.SO_PATH = os.path.join(ROOT, "cuda_voxelizer_b200", os.environ["VOXB200_SO"])
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
name, G = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("icosphere:708:1024", 2048)
v, f = cases.mesh(name)
d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
grid = vb.grid_from_verts(v, G, len(f))
table = torch.empty(vb.table_bytes(G) // 4, dtype=torch.int32, device="cuda")
m = vb.Mesh(grid, tris=d)
for _ in range(int(os.environ.get("REPS", "5"))):
    m.voxelize(table=table)
torch.cuda.synchronize()
m.close()
