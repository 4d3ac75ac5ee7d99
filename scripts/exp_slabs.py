"""Experiment: config 4 on ONE GPU as S z-slabs processed one after the other (triangles routed per slab), to see
whether slab-local atomics (slab <= L2) remove the DRAM read-modify-write traffic.  VOXB200_SO picks the library build."""
import copy, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuda_voxelizer_b200 import _lib
if os.environ.get("VOXB200_SO"):
    _lib.SO_PATH = os.path.join(ROOT, "cuda_voxelizer_b200", os.environ["VOXB200_SO"])
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
name, G = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("icosphere:708:1024", 2048)
slab_counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 8, 16, 32]
v, f = cases.mesh(name)
soup = np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))
if os.environ.get("ZSORT"):
    # experiment: triangles ordered by the z of their lowest vertex (what an upload-time sort would produce)
    soup = np.ascontiguousarray(soup[np.argsort(soup[:, 2::3].min(axis=1), kind="stable")])
d = torch.from_numpy(soup).cuda()
T = len(f)
grid = vb.grid_from_verts(v, G, T)
words = vb.table_bytes(G) // 4
table = torch.empty(words, dtype=torch.int32, device="cuda")
ref = torch.empty(words, dtype=torch.int32, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: vb.voxelize(grid, d, table=ref))
vb.set_profiling(True); vb.voxelize(grid, d, table=ref); torch.cuda.synchronize(); ph = vb.phase_ms(0); vb.set_profiling(False)
w = (torch.arange(words, device="cuda", dtype=torch.int64) % 65521) + 1
chk = int((ref.long() * w).sum().item()) & 0xffffffffffffffff
del w
print("%s lib=%s direct: %.4f ms  phases %s  checksum %016x" % (name, os.path.basename(_lib.SO_PATH), ms, ["%.4f" % x for x in ph], chk), flush=True)
for S in slab_counts:
    if S == 1: continue
    regions = [vb.partition(G, False, r, S)[0] for r in range(S)]
    out = torch.empty(int(T * 1.25) * 9, dtype=torch.float32, device="cuda")
    counts = vb.route_triangles_multi(grid, d, regions, out)
    t_route = timeit(lambda: vb.route_triangles_multi(grid, d, regions, out), n=3)
    offs = np.concatenate([[0], np.cumsum(counts)])
    segs, grids, tabs = [], [], []
    wps = words // S
    for r in range(S):
        segs.append(out[int(offs[r]) * 9:int(offs[r + 1]) * 9])
        g = vb.Grid.from_buffer_copy(bytes(grid)); g.n_triangles = counts[r]; grids.append(g)
        tabs.append(table[r * wps:(r + 1) * wps])
    def run():
        for r in range(S):
            vb.voxelize(grids[r], segs[r], table=tabs[r], region=regions[r])
    ms = timeit(run)
    vb.set_profiling(True); run(); torch.cuda.synchronize()
    ph = np.array([vb.phase_ms(i) for i in range(S)]).sum(axis=0); vb.set_profiling(False)
    same = bool(torch.equal(table, ref))
    print("S=%2d slabs: %.4f ms total (sum of phases: zero %.4f tri %.4f coop %.4f)  routed %d (+%.2f%%) route %.3f ms  identical=%s"
          % (S, ms, ph[0], ph[1], ph[2], sum(counts), 100.0 * (sum(counts) - T) / T, t_route, same), flush=True)
