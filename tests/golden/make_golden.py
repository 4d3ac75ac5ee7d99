"""Generates tests/golden/{bunny.npz,golden.json,tables_small.npz} from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

Every number written here is an output of the reference's unmodified src/cpu_voxelizer.cpp,
compiled by oracle/Makefile into oracle/_ref/libvoxref.so and driven through oracle/ref_driver.cpp
(bbox -> createMeshBBCube -> voxinfo -> cpu_voxelize_mesh{,_solid}), with OMP_NUM_THREADS=1 (the
result is thread-count independent — OR/XOR commute — and one thread is the fast setting).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from cuda_voxelizer_b200 import meshio  # noqa: E402

REF_BUNNY = "/root/reference/test_models/bunny.OBJ"


def main():
    only_missing = "--update" in sys.argv
    verts, faces = meshio.read_obj(REF_BUNNY)
    assert verts.shape == (2557, 3) and faces.shape == (5110, 3)
    np.savez_compressed(os.path.join(HERE, "bunny.npz"), verts=verts, faces=faces)

    import cases
    out_path = os.path.join(HERE, "golden.json")
    golden = json.load(open(out_path)) if (only_missing and os.path.exists(out_path)) else {}
    small_tables = {}
    meshes = {}
    for name, g, solid, morton in cases.GOLDEN_CASES:
        key = cases.case_key(name, g, solid, morton)
        if key in golden and only_missing:
            continue
        if name not in meshes:
            meshes.clear()
            meshes[name] = cases.mesh(name)
        v, f = meshes[name]
        t0 = time.time()
        table, ms = oracle.ref_voxelize(v, f, g, solid=solid, morton=morton, threads=1, return_ms=True)
        mn, mx, unit = oracle.ref_voxinfo(v, g, len(f))
        golden[key] = {
            "n_verts": int(len(v)), "n_tris": int(len(f)),
            "popcount": oracle.popcount(table), "fnv1a64": "%016x" % oracle.fnv1a64(table),
            "bbox_min": [float(x) for x in mn], "bbox_max": [float(x) for x in mx], "unit": [float(x) for x in unit],
            "ref_ms_1thread": round(ms, 2),
        }
        if g <= 64:
            small_tables[key] = table
        print("%-55s popcount %12d  %s  ref %.1f ms (total %.1fs)" % (key, golden[key]["popcount"], golden[key]["fnv1a64"], ms, time.time() - t0), flush=True)
        del table
        json.dump(golden, open(out_path, "w"), indent=1, sort_keys=True)
    if small_tables:
        np.savez_compressed(os.path.join(HERE, "tables_small.npz"), **small_tables)
    print("voxinfo layout:", oracle.ref_voxinfo_layout())
    make_writer_goldens(verts, faces)


def make_writer_goldens(verts, faces):
    """Golden OUTPUT FILES from the reference's own writers (util_io.cpp), bunny surface @32 (morton table for -o morton)."""
    import shutil
    import tempfile
    g = 32
    out_dir = os.path.join(HERE, "io")
    os.makedirs(out_dir, exist_ok=True)
    mn, mx, unit = oracle.ref_voxinfo(verts, g, len(faces))
    lin = oracle.ref_voxelize(verts, faces, g, threads=1)
    mor = oracle.ref_voxelize(verts, faces, g, morton=True, threads=1)
    tmp = tempfile.mkdtemp()
    base = os.path.join(tmp, "bunny.OBJ")
    index = {}
    for fmt, table, produced in (("binvox", lin, "bunny.OBJ_%d.binvox" % g), ("morton", mor, "bunny.OBJ.bin"),
                                 ("obj_points", lin, "bunny.OBJ_%d_pointcloud.obj" % g), ("obj", lin, "bunny.OBJ_%d_voxels.obj" % g),
                                 ("vox", lin, "bunny.OBJ_%d.vox" % g)):
        oracle.ref_write(fmt, table, g, mn, mx, len(faces), base)
        data = open(os.path.join(tmp, produced), "rb").read()
        index[fmt] = {"file": produced, "bytes": len(data), "fnv1a64": "%016x" % oracle.fnv1a64(np.frombuffer(data, np.uint8))}
        if fmt in ("binvox", "vox"):
            shutil.copy(os.path.join(tmp, produced), os.path.join(out_dir, produced))
        print("writer golden %-10s %-32s %8d bytes %s" % (fmt, produced, len(data), index[fmt]["fnv1a64"]))
    json.dump({"gridsize": g, "files": index}, open(os.path.join(out_dir, "index.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp)


def make_binvox_goldens():
    """Size + FNV-1a-64 of the reference writer's binvox FILES at grid sizes the device encoder covers (voxb200_binvox_rle)."""
    import shutil
    import tempfile
    import cases
    v, f = cases.mesh("bunny")
    out = {}
    tmp = tempfile.mkdtemp()
    for g, solid in ((256, 0), (256, 1), (512, 0)):
        mn, mx, unit = oracle.ref_voxinfo(v, g, len(f))
        table = oracle.ref_voxelize(v, f, g, solid=bool(solid), threads=1)
        base = os.path.join(tmp, "bunny.OBJ")
        oracle.ref_write("binvox", table, g, mn, mx, len(f), base)
        data = np.fromfile(base + "_%d.binvox" % g, np.uint8)
        header_end = data.tobytes().index(b"data\n") + 5
        key = "bunny|%d|%s" % (g, "solid" if solid else "surface")
        out[key] = {"bytes": int(len(data)), "header_bytes": int(header_end), "fnv1a64": "%016x" % oracle.fnv1a64(data),
                    "payload_fnv1a64": "%016x" % oracle.fnv1a64(data[header_end:])}
        print("binvox golden", key, out[key])
    json.dump(out, open(os.path.join(HERE, "io", "binvox_large.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    if "--binvox" in sys.argv:
        make_binvox_goldens()
    else:
        main()
        make_binvox_goldens()
