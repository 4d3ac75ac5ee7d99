"""GPU suite (-m gpu) of the multi-GPU C-ABI entry points: voxb200_voxelize_host_multi (one process, one host thread per device)
and voxb200_gather_slabs.  Runs with however many GPUs the box has (1 on the round-end box, more under `gpurun --gpus N`); the
table must equal the single-GPU table / the reference golden for every device count."""
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)
    return vb


def _device_counts(vb):
    n = vb.device_count()
    return sorted({1, min(2, n), min(4, n), n})


@pytest.mark.parametrize("name,g,solid,morton", [("icosphere:64:128", 256, 0, 0), ("bunny", 256, 0, 0), ("bunny", 128, 1, 0),
                                                  ("bunny", 64, 0, 1), ("bunny", 256, 1, 1), ("torus:100:50:256", 256, 0, 0)])
def test_host_multi_matches_golden(vb, golden, name, g, solid, morton):
    want = golden[cases.case_key(name, g, solid, morton)]
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    hv = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    hf = torch.from_numpy(np.ascontiguousarray(f)).pin_memory()
    for n in _device_counts(vb):
        if morton and n not in (1, 2, 4, 8):
            continue
        out = torch.full((vb.table_bytes(g) // 4,), -1, dtype=torch.int32).pin_memory()
        for _ in range(2):                      # the second call reuses every buffer and re-prepares the meshes
            table, ms = vb.voxelize_host_multi(grid, hv, hf, out, solid=bool(solid), morton=bool(morton), n_devices=n)
        host = out.numpy().view(np.uint32)
        assert oracle.popcount(host) == want["popcount"], "n_devices=%d" % n
        assert "%016x" % oracle.fnv1a64(host) == want["fnv1a64"], "n_devices=%d" % n
        assert ms[7] == n and ms[5] > 0.0


def test_host_multi_from_pageable_memory(vb, golden):
    name, g = "icosphere:16:64", 128
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    table, ms = vb.voxelize_host_multi(grid, np.ascontiguousarray(v), np.ascontiguousarray(f), n_devices=vb.device_count())
    want = golden[cases.case_key(name, g, 0, 0)]
    assert "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]


def test_host_multi_rejects_bad_device_lists(vb):
    v, f = cases.mesh("bunny")
    grid = vb.grid_from_verts(v, 64, len(f))
    with pytest.raises(vb.VoxError):
        vb.voxelize_host_multi(grid, v, f, devices=[0, 0])
    with pytest.raises(vb.VoxError):
        vb.voxelize_host_multi(grid, v, f, devices=[vb.device_count()])


def test_gather_slabs(vb):
    """Slabs voxelized per region (on as many devices as there are) gathered into one table on device 0 == the 1-GPU table."""
    name, g, n_parts = "icosphere:64:128", 256, 4
    v, f = cases.mesh(name)
    soup = oracle.soup(v, f)
    grid = vb.grid_from_verts(v, g, len(f))
    n_dev = vb.device_count()
    slabs, devs = [], []
    for p in range(n_parts):
        dev = p % n_dev
        torch.cuda.set_device(dev)
        vb.init(dev)
        region, nbytes = vb.partition(g, False, p, n_parts)
        slabs.append(vb.voxelize(grid, torch.from_numpy(soup).cuda(dev), region=region))
        devs.append(dev)
        torch.cuda.synchronize(dev)
    torch.cuda.set_device(0)
    vb.init(0)
    full = vb.voxelize(grid, torch.from_numpy(soup).cuda(0))
    out = torch.empty_like(full)
    vb.gather_slabs(slabs, devs, out, 0)
    torch.cuda.synchronize()
    assert torch.equal(out, full)


def test_c_caller_of_host_multi(vb, golden, tmp_path):
    """The entry point from plain C (tests/c/test_multi.c), compiled here with gcc against include/voxb200.h."""
    exe = tmp_path / "test_multi"
    lib_dir = os.path.join(ROOT, "cuda_voxelizer_b200")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "test_multi.c"),
                           "-L", lib_dir, "-lvoxb200", "-Wl,-rpath," + lib_dir, "-o", str(exe)])
    name, g = "icosphere:64:128", 256
    v, f = cases.mesh(name)
    mesh_bin = tmp_path / "mesh.bin"
    with open(mesh_bin, "wb") as fh:
        fh.write(np.array([len(v), len(f)], np.uint64).tobytes())
        fh.write(np.ascontiguousarray(v, np.float32).tobytes())
        fh.write(np.ascontiguousarray(f, np.int32).tobytes())
    for flags, key in ((0, cases.case_key(name, g, 0, 0)), (8, cases.case_key(name, g, 1, 0))):
        out = subprocess.run([str(exe), str(mesh_bin), str(g), str(flags), "0"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        words = out.stdout.split()
        assert words[1] == golden[key]["fnv1a64"] and int(words[3]) == golden[key]["popcount"], out.stdout
