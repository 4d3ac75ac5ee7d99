"""Timing of the secondary modes (morton, solid+morton, SoA4) on bunny and the 1M sphere."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, G in (("bunny", 1024), ("icosphere:224:512", 1024), ("icosphere:708:1024", 2048)):
    v, f = cases.mesh(name)
    soup = np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))
    d = torch.from_numpy(soup).cuda()
    grid = vb.grid_from_verts(v, G, len(f))
    table = torch.empty(vb.table_bytes(G) // 4, dtype=torch.int32, device="cuda")
    for solid in (0, 1):
        if solid and G > 1024: continue
        for morton in (0, 1):
            fn = vb.voxelize_solid if solid else vb.voxelize
            ms = timeit(lambda: fn(grid, d, table=table, morton=bool(morton)))
            print("%-22s G=%d %s %s : %.3f ms" % (name, G, "solid" if solid else "surface", "morton" if morton else "linear", ms))
