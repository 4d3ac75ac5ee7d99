"""Experiment: voxb200_voxelize_host_multi on all GPUs of the box (config 4), phases and wall time; environment picks the read-back options."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_voxelizer_b200 as vb
import cases
vb.init(0)
v, f = cases.mesh("icosphere:708:1024")
G = 2048
grid = vb.grid_from_verts(v, G, len(f))
hv = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
hf = torch.from_numpy(np.ascontiguousarray(f)).pin_memory()
out = torch.empty(vb.table_bytes(G) // 4, dtype=torch.int32).pin_memory()
tag = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("VOXB200_"))
for n in [int(x) for x in os.environ.get("NDEV", "1,2,4,8").split(",")]:
    if n > vb.device_count():
        continue
    for _ in range(2):
        vb.voxelize_host_multi(grid, hv, hf, out, n_devices=n)
    acc = np.zeros(8); K = 6
    t0 = time.perf_counter()
    for _ in range(K):
        _, tm = vb.voxelize_host_multi(grid, hv, hf, out, n_devices=n)
        acc += np.array(tm)
    wall = (time.perf_counter() - t0) * 1e3 / K
    acc /= K
    print("[%s] n=%d wall %.2f ms | h2d %.2f gather %.2f prepare %.2f voxelize %.2f back %.2f total %.2f" % (tag, n, wall, *acc[:6]), flush=True)
