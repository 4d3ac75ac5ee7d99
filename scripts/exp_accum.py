"""Experiment: per-triangle kernel time with and without the zero-fill right before it (L2 state)."""
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
table = torch.zeros(vb.table_bytes(G) // 4, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(acc, n=10, flush_l2=False):
    vb.set_profiling(True)
    for _ in range(n):
        if flush_l2: flush.fill_(1)
        vb.voxelize(grid, d, table=table, accumulate=acc)
    torch.cuda.synchronize()
    ph = np.array([vb.phase_ms(i) for i in range(n)]).mean(0)
    vb.set_profiling(False)
    return ph
print("zero+tri      :", run(False))
print("accumulate    :", run(True))
print("accum+flushL2 :", run(True, flush_l2=True))
