"""Experiment: the prepared-mesh tile path on the z-slabs of an 8-way split, one GPU (what each rank of the N = 8 bench runs).
VOXB200_SO picks the library build (tile geometry variants)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuda_voxelizer_b200 import _lib
if os.environ.get("VOXB200_SO"):
    _lib.SO_PATH = os.path.join(ROOT, "cuda_voxelizer_b200", os.environ["VOXB200_SO"])
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
N = int(os.environ.get("PARTS", "8"))
v, f = cases.mesh("icosphere:708:1024")
G = 2048
d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
grid = vb.grid_from_verts(v, G, len(f))
full = vb.voxelize(grid, d)
torch.cuda.synchronize()
words = full.numel()
out = []
for r in range(N):
    region, nbytes = vb.partition(G, False, r, N)
    routed, n = vb.route_triangles(grid, d, region)
    import copy
    g2 = copy.copy(grid); g2.n_triangles = n
    tris = torch.empty(0)
    m = vb.Mesh(g2, tris=routed, region=region)
    table = torch.empty(nbytes // 4, dtype=torch.int32, device="cuda")
    for _ in range(3):
        m.voxelize(table=table)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.voxelize(table=table)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    same = bool(torch.equal(table, full[r * (words // N):(r + 1) * (words // N)]))
    info = m.info()
    out.append(ms)
    print("slab %d/%d: %8d tris, %5d work tiles, %8d records: %.4f ms  identical=%s" % (r, N, n, info["work_tiles"], info["instances"], ms, same), flush=True)
    m.close()
print("lib=%s max %.4f ms, mean %.4f ms (1-GPU whole mesh / %d would be %.4f)" % (os.path.basename(_lib.SO_PATH), max(out), sum(out) / len(out), N, 0.4253 / N))
