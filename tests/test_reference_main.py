"""GPU suite: the link-level drop-in.  oracle/_ref/cuda_voxelizer_refmain is the reference's UNMODIFIED src/main.cpp,
src/util_io.cpp and src/cpu_voxelizer.cpp (compiled from /root/reference by oracle/Makefile, trimesh2 replaced by the
test shim) linked against OUR libvoxb200.so, which supplies voxelize(), voxelize_solid() and initCuda() in place of the
reference's voxelize.cu / voxelize_solid.cu / util_cuda.cpp.  Running it exercises the reference's own caller:
cudaMallocManaged triangles and table (main.cpp:66,214), voxelize(info, tris, vtable, morton), the reference writers."""
import json
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFMAIN = os.path.join(ROOT, "oracle", "_ref", "cuda_voxelizer_refmain")
IO_DIR = os.path.join(ROOT, "tests", "golden", "io")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REFMAIN), reason="oracle/_ref/cuda_voxelizer_refmain not built")]


@pytest.fixture(scope="module")
def obj_path(tmp_path_factory):
    from cuda_voxelizer_b200 import meshio
    d = tmp_path_factory.mktemp("refmain")
    v, f = cases.mesh("bunny")
    p = str(d / "bunny.OBJ")
    meshio.write_obj(p, v, f)
    return p


def test_reference_main_runs_on_our_library(obj_path, golden):
    idx = json.load(open(os.path.join(IO_DIR, "index.json")))
    g = idx["gridsize"]
    r = subprocess.run([REFMAIN, "-f", obj_path, "-s", str(g), "-o", "binvox"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "## GPU VOXELISATION" in r.stdout and "[Perf] Voxelization GPU time:" in r.stdout      # the GPU branch, through voxelize()
    produced = os.path.join(os.path.dirname(obj_path), "bunny.OBJ_%d.binvox" % g)
    assert open(produced, "rb").read() == open(os.path.join(IO_DIR, "bunny.OBJ_%d.binvox" % g), "rb").read()
    for flags, key in ((["-o", "morton"], ("bunny", 256, 0, 1)), (["-solid", "-o", "morton"], ("bunny", 256, 1, 1))):
        r = subprocess.run([REFMAIN, "-f", obj_path] + flags, capture_output=True, text=True, timeout=300)       # default -s 256
        assert r.returncode == 0, r.stdout + r.stderr
        table = np.fromfile(obj_path + ".bin", np.uint32)
        want = golden[cases.case_key(*key)]
        assert oracle.popcount(table) == want["popcount"] and "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]
