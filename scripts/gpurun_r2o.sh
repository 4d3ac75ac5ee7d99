#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_binvox.py tests/test_cli.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2o_pytest_binvox.txt
# timing of the encoder on the config-4 table
timeout 600 python - <<'P' 2>&1 | tee gpurun_out/r2o_binvox_time.log
import sys, time, numpy as np, torch
sys.path.insert(0, "tests")
import cuda_voxelizer_b200 as vb, cases
vb.init(0)
v, f = cases.mesh("icosphere:708:1024")
d = torch.from_numpy(np.ascontiguousarray(v[f.reshape(-1)].reshape(-1, 9))).cuda()
grid = vb.grid_from_verts(v, 2048, len(f))
t = vb.voxelize(grid, d)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); p = vb.binvox_rle(t, 2048); dt = time.perf_counter() - t0
    print("binvox payload of config 4 @2048^3: %d bytes in %.2f ms (incl. allocations and the D2H of the payload)" % (len(p), dt * 1e3))
P
