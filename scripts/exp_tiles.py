"""Experiment: the tile-owner schedule of a prepared mesh (voxb200_mesh_*) against the one-shot kernels, config 4 by default.
Prints per-path device time (CUDA events, after warm-up), the prepare cost and the handle's plan; checks the table hash.
VOXB200_SO picks the library build."""
import os, sys, time, numpy as np, torch
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
soup = np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))
d = torch.from_numpy(soup).cuda()
T = len(f)
grid = vb.grid_from_verts(v, G, T)
words = vb.table_bytes(G) // 4
table = torch.empty(words, dtype=torch.int32, device="cuda")
ref = torch.empty(words, dtype=torch.int32, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lib = os.path.basename(_lib.SO_PATH)
if not os.environ.get("SKIP_ONESHOT"):
    ms = timeit(lambda: vb.voxelize(grid, d, table=ref))
    print("%s @%d lib=%s one-shot, caller's order: %.4f ms" % (name, G, lib, ms), flush=True)
    ds = vb.sort_triangles(grid, d)
    ms = timeit(lambda: vb.voxelize(grid, ds, table=ref))
    print("one-shot, z-layer order: %.4f ms" % ms, flush=True)
    ds.close()
else:
    vb.voxelize(grid, d, table=ref)
torch.cuda.synchronize()
t0 = time.perf_counter()
m = vb.Mesh(grid, tris=d)
t_create = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
for _ in range(5):
    m.update(tris=d)
t_update = (time.perf_counter() - t0) * 1e3 / 5
print("mesh create %.2f ms wall (first, with allocations); update (re-prepare, buffers reused) %.3f ms wall; plan %s" % (t_create, t_update, m.info()), flush=True)
ms = timeit(lambda: m.voxelize(table=table))
same = bool(torch.equal(table, ref))
print("prepared mesh (tile schedule): %.4f ms  identical=%s  -> %.0f Mtri/s, step roofline %.3f of 6559 GB/s"
      % (ms, same, T / ms / 1e3, (36 * T + words * 4) / (ms * 1e-3) / 1e9 / 6559.4), flush=True)
vb.set_profiling(True)
for _ in range(4):
    m.voxelize(table=table)
torch.cuda.synchronize()
print("phases (tile kernel, side path): %s" % [["%.4f" % x for x in vb.phase_ms(i)[1:3]] for i in range(4)], flush=True)
vb.set_profiling(False)
vi = torch.from_numpy(np.ascontiguousarray(v)).cuda()
fi = torch.from_numpy(np.ascontiguousarray(f)).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
mi = vb.Mesh(grid, verts=vi, faces=fi)
t_ci = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
for _ in range(5):
    mi.update(verts=vi, faces=fi)
t_ui = (time.perf_counter() - t0) * 1e3 / 5
mi.voxelize(table=table)
print("indexed create %.2f ms, update %.3f ms wall; identical=%s" % (t_ci, t_ui, bool(torch.equal(table, ref))), flush=True)
