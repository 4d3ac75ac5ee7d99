"""The reference's own GPU kernels (oracle/_ref/libvoxref_gpu.so: unmodified voxelize.cu / voxelize_solid.cu built for sm_100a)
timed on this GPU for BASELINE configs 2, 3, 4 — triangles and table device-resident, CUDA events around the reference's
voxelize() / voxelize_solid().  Test/bench infrastructure: prints one JSON line per config."""
import json, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import cases
import cuda_voxelizer_b200 as vb
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
for name, G, solid in (("bunny", 1024, 0), ("icosphere:224:512", 1024, 1), ("icosphere:708:1024", 2048, 0)):
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, G, len(f))
    soup = oracle.soup(v, f)
    ms, table = oracle.ref_gpu_run(list(grid.bbox_min), list(grid.bbox_max), G, soup, solid=bool(solid), warmup=2, reps=5, want_table=True)
    gold = golden[cases.case_key(name, G, solid, 0)]
    print(json.dumps({"mesh": name, "G": G, "solid": solid, "triangles": len(f), **{k: round(x, 4) for k, x in ms.items()},
                      "popcount": oracle.popcount(table), "cpu_reference_popcount": gold["popcount"]}), flush=True)
