"""Experiment: does a 1 GiB zero-fill overlap with the per-triangle kernel when both run at once (two streams, two tables)?
Tells whether hiding the zero-fill behind the arithmetic is possible in principle on this part."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
name, G = "icosphere:708:1024", 2048
v, f = cases.mesh(name)
d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
grid = vb.grid_from_verts(v, G, len(f))
words = vb.table_bytes(G) // 4
tA = torch.zeros(words, dtype=torch.int32, device="cuda")
tB = torch.empty(words, dtype=torch.int32, device="cuda")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def run(which, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        sA.wait_stream(torch.cuda.current_stream()); sB.wait_stream(torch.cuda.current_stream())
        if which in ("tri", "both"):
            vb.voxelize(grid, d, table=tA, accumulate=True, stream=sA)      # no zero-fill inside: OR into tA
        if which in ("zero", "both"):
            with torch.cuda.stream(sB):
                tB.zero_()
        torch.cuda.current_stream().wait_stream(sA); torch.cuda.current_stream().wait_stream(sB)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for w in ("tri", "zero", "both", "tri", "zero", "both"):
    run(w, 3)
    print("%-5s %.4f ms" % (w, run(w)), flush=True)
