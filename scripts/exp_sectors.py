"""Experiment: how many distinct 32-byte sectors / 128-byte lines / table words does config 4 touch?"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_voxelizer_b200 as vb
from cuda_voxelizer_b200 import meshgen
vb.init(0)
v, f = meshgen.icosphere(708, radius=1024.0)
soup = np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))
d = torch.from_numpy(soup).cuda()
G = 2048
grid = vb.grid_from_verts(v, G, len(f))
t = vb.voxelize(grid, d)
torch.cuda.synchronize()
nz = t != 0
print("nonzero words   :", int(nz.sum()))
print("nonzero sectors :", int(nz.view(-1, 8).any(dim=1).sum()), "(32 B)")
print("nonzero lines   :", int(nz.view(-1, 32).any(dim=1).sum()), "(128 B)")
print("nonzero rows    :", int(nz.view(-1, 64).any(dim=1).sum()), "(256 B rows)")
