"""Timing of the cooperative (large-triangle) regimes for A/B runs of the coop kernel.  VOXB200_SO picks the library build."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuda_voxelizer_b200 import _lib
if os.environ.get("VOXB200_SO"):
    _lib.SO_PATH = os.path.join(ROOT, "cuda_voxelizer_b200", os.environ["VOXB200_SO"])
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, G in (("bunny", 1024), ("bunny", 2048), ("icosphere:59:512", 1024), ("icosphere:59:1024", 2048), ("icosphere:224:2048", 4096), ("icosphere:224:512", 2048), ("soup:mixed:20000:3:1.0", 1024)):
    v, f = cases.mesh(name)
    d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
    grid = vb.grid_from_verts(v, G, len(f))
    table = torch.empty(vb.table_bytes(G) // 4, dtype=torch.int32, device="cuda")
    for morton in (0,):
        if morton and G > 2048: continue
        ms = timeit(lambda: vb.voxelize(grid, d, table=table, morton=bool(morton)))
        vb.set_profiling(True); vb.voxelize(grid, d, table=table, morton=bool(morton)); torch.cuda.synchronize(); ph = vb.phase_ms(0); vb.set_profiling(False)
        chk = int(table.long().sum().item()) & 0xffffffffffff
        print("%-26s G=%d %s: %.4f ms (coop %.4f) lib=%s chk %012x" % (name, G, "morton" if morton else "linear", ms, ph[2], os.path.basename(_lib.SO_PATH), chk), flush=True)
