"""Experiment: the read-back of the config-4 table (voxb200_download_table) against thread count and mode."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
name, G = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("icosphere:708:1024", 2048)
v, f = cases.mesh(name)
d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
grid = vb.grid_from_verts(v, G, len(f))
words = vb.table_bytes(G) // 4
table = torch.empty(words, dtype=torch.int32, device="cuda")
vb.voxelize(grid, d, table=table)
torch.cuda.synchronize()
host = torch.empty(words, dtype=torch.int32).pin_memory()
ref = table.cpu()
for mode, threads in [("dense", 0), ("sparse", 1), ("sparse", 2), ("sparse", 4), ("sparse", 6), ("sparse", 8), ("sparse", 12), ("sparse", 16)]:
    vb.set_readback_mode(mode); vb.set_host_threads(threads)
    best = 1e9
    for _ in range(4):
        host.fill_(-1)
        t0 = time.perf_counter()
        vb.download_table(table, host)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    print("%s threads=%d: %.3f ms  identical=%s" % (mode, threads, best, bool(torch.equal(host, ref))), flush=True)
