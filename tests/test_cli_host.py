"""CPU suite: the CLI's host side without a GPU — flag handling, the OBJ/PLY loader + voxinfo print, and (through the
`--from-table` test hook, which feeds the writers a table produced by the ORACLE) the five writers, byte-compared with
the golden files produced by the reference's own writers (tests/golden/io)."""
import json
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda_voxelizer_b200", "bin", "cuda_voxelizer")
IO_DIR = os.path.join(ROOT, "tests", "golden", "io")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    from cuda_voxelizer_b200 import _lib, meshio
    if not os.path.exists(CLI):
        _lib.build()
    d = tmp_path_factory.mktemp("clihost")
    v, f = cases.mesh("bunny")
    obj = str(d / "bunny.OBJ")
    meshio.write_obj(obj, v, f)
    g = json.load(open(os.path.join(IO_DIR, "index.json")))["gridsize"]
    mn, mx, unit = oracle.voxinfo(v, g)
    soup = oracle.soup(v, f)
    lin, mor = str(d / "lin.tbl"), str(d / "mor.tbl")
    oracle.surface(soup, mn, unit, g).tofile(lin)
    oracle.surface(soup, mn, unit, g, morton=True).tofile(mor)
    return {"obj": obj, "g": g, "lin": lin, "mor": mor, "dir": str(d)}


def _run(args):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=120)


def _fnv(path):
    return "%016x" % oracle.fnv1a64(np.frombuffer(open(path, "rb").read(), np.uint8))


def test_writers_match_reference_writers_without_gpu(work):
    idx = json.load(open(os.path.join(IO_DIR, "index.json")))
    for fmt, key, table in (("binvox", "binvox", "lin"), ("morton", "morton", "mor"), ("obj_points", "obj_points", "lin"), ("obj", "obj", "lin")):
        r = _run(["-f", work["obj"], "-s", str(work["g"]), "-o", fmt, "--from-table", work[table]])
        assert r.returncode == 0, r.stdout + r.stderr
        produced = os.path.join(work["dir"], idx["files"][key]["file"])
        assert os.path.getsize(produced) == idx["files"][key]["bytes"], fmt
        assert _fnv(produced) == idx["files"][key]["fnv1a64"], fmt
    r = _run(["-f", work["obj"], "-s", str(work["g"]), "-o", "vox", "--from-table", work["lin"]])
    assert r.returncode == 0 and os.path.getsize(os.path.join(work["dir"], "bunny.OBJ_%d.vox" % work["g"])) > 1000


def test_loader_and_voxinfo_print(work, golden):
    r = _run(["-f", work["obj"], "-s", "64", "-o", "binvox", "--from-table", "/nonexistent"])
    assert "[Mesh] Number of triangles: 5110" in r.stdout and "[Mesh] Number of vertices: 2557" in r.stdout
    assert "[Voxelization] Bounding Box: (-2.966815,0.033993,-2.477184)-(1.915775,4.916584,2.405406)" in r.stdout
    assert "Unit length: x: 0.076290 y: 0.076290 z: 0.076290" in r.stdout
    assert r.returncode == 1 and "cannot read" in r.stdout


def test_flags_without_gpu(work):
    assert _run([]).returncode == 0 and _run(["-h"]).returncode == 0
    r = _run(["-s", "64"])
    assert r.returncode == 1 and "didn't specify a file" in r.stdout
    r = _run(["-f", work["obj"], "-o", "nonsense"])
    assert r.returncode == 1 and "Unrecognized output format" in r.stdout
    r = _run(["-f", work["obj"], "-s", "64", "-cpu"])
    assert r.returncode == 1 and "no CPU voxelization path" in r.stdout
