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


def _dump(path, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([CLI, "-f", path, "-s", "64", "--dump-mesh", out], capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    nv, nf = np.frombuffer(raw[:16], np.uint64)
    v = np.frombuffer(raw[16:16 + 12 * int(nv)], np.float32).reshape(-1, 3)
    f = np.frombuffer(raw[16 + 12 * int(nv):], np.int32).reshape(-1, 3)
    assert len(f) == nf
    return v, f


def test_parallel_obj_loader_matches_the_line_by_line_parser(work, tmp_path):
    """The memory-mapped, multi-threaded OBJ parser (ingest) against the fgets/strtof one it replaces, bit for
    bit: number formats, '+' signs, CRLF, comments, vn/vt lines, a/t/n index forms, polygons, relative indices across
    chunk boundaries, at several thread counts (the file is > 1 MB so that it really is cut into chunks)."""
    rng = np.random.default_rng(7)
    lines = ["# header comment", "mtllib x.mtl", "o thing"]
    n_blocks = 5000
    for b in range(n_blocks):
        for k in range(4):
            x, y, z = rng.normal(0, 10, 3)
            fmt = ("v %.9g %.9g %.9g", "v %+.7e %.3f %.12g 1.0", "v\t%f  %g %.1f 0.5 0.25 0.125", "  v %.9g %.9g %.9g")[(b + k) % 4]
            lines.append(fmt % (x, y, z))
        lines.append("vn 0 0 1")
        lines.append("vt 0.5 0.5")
        base = 4 * b + 1
        form = b % 5
        if form == 0:
            lines.append("f %d %d %d" % (base, base + 1, base + 2))
        elif form == 1:
            lines.append("f %d/1/1 %d/1/1 %d/1/1 %d/1/1" % (base, base + 1, base + 2, base + 3))        # a quad: fanned
        elif form == 2:
            lines.append("f %d//1 %d//1 %d//1" % (base + 3, base + 2, base + 1))
        elif form == 3:
            lines.append("f -4 -3 -2")                                                                  # relative to the vertices read so far
        else:
            lines.append("f -1/1 -2/1 -3/1 -4/1")
        if b % 97 == 0:
            lines.append("")
            lines.append("# " + "x" * 3000)                                                             # a long comment line
    text = "\n".join(lines) + "\n"
    assert len(text) > (1 << 20)
    unix, dos = str(tmp_path / "mix.obj"), str(tmp_path / "mix_crlf.obj")
    open(unix, "w").write(text)
    open(dos, "w", newline="").write(text.replace("\n", "\r\n"))
    ref_v, ref_f = _dump(unix, str(tmp_path / "ref.bin"), {"VOXCLI_SERIAL_LOADER": "1"})
    assert len(ref_v) == 4 * n_blocks and ref_f.min() >= 0 and ref_f.max() < len(ref_v)
    for path in (unix, dos):
        for threads in ("1", "2", "3", "8", "31"):
            v, f = _dump(path, str(tmp_path / "par.bin"), {"VOXCLI_LOADER_THREADS": threads})
            assert v.tobytes() == ref_v.tobytes() and f.tobytes() == ref_f.tobytes(), (path, threads)
    # and the bundled bunny, against the arrays the file was written from
    v, f = _dump(work["obj"], str(tmp_path / "bunny.bin"))
    bv, bf = cases.mesh("bunny")
    assert np.array_equal(v, bv) and np.array_equal(f, bf)


def _tess_quad(v, q):
    """trimesh2's tess() rule for a quad: split along the shorter diagonal."""
    d02 = np.float32(((v[q[0]] - v[q[2]]) ** 2).sum(dtype=np.float32))
    d13 = np.float32(((v[q[1]] - v[q[3]]) ** 2).sum(dtype=np.float32))
    i = 0 if d02 < d13 else 1
    return [[q[i], q[(i + 1) % 4], q[(i + 2) % 4]], [q[i], q[(i + 2) % 4], q[(i + 3) % 4]]]


def test_quads_are_split_along_the_shorter_diagonal(tmp_path):
    """OBJ and PLY polygons: triangles as they are, quads along the shorter diagonal, larger polygons as a fan (trimesh2 tess())."""
    v = np.array([[0, 0, 0], [4, 0, 0], [4, 1, 0], [0, 1, 0],          # d02 == d13: corner 1
                  [0, 0, 1], [1, 0, 1], [5, 3, 1], [0, 1, 1],          # d13 shorter: corner 1
                  [0, 0, 2], [1, 0, 2], [1, 1, 2], [-5, 4, 2],         # d02 shorter: corner 0
                  [9, 9, 9], [8, 9, 9], [8, 8, 9], [9, 7, 9], [9.5, 8, 9]], np.float32)
    polys = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15, 16], [0, 4, 8]]
    want = []
    for p in polys:
        if len(p) == 4:
            want += _tess_quad(v, p)
        else:
            want += [[p[0], p[k], p[k + 1]] for k in range(1, len(p) - 1)]
    want = np.array(want, np.int32)
    obj = tmp_path / "quads.obj"
    with open(obj, "w") as fh:
        for p in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(p))
        for p in polys:
            fh.write("f " + " ".join(str(i + 1) for i in p) + "\n")
    for env in ({"VOXCLI_SERIAL_LOADER": "1"}, {}):
        gv, gf = _dump(str(obj), str(tmp_path / "q.bin"), env)
        assert np.array_equal(gv, v) and np.array_equal(gf, want), env
    ply = tmp_path / "quads.ply"
    with open(ply, "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(v), len(polys)))
        for p in v:
            fh.write("%.9g %.9g %.9g\n" % tuple(p))
        for p in polys:
            fh.write("%d %s\n" % (len(p), " ".join(map(str, p))))
    gv, gf = _dump(str(ply), str(tmp_path / "q.bin"))
    assert np.array_equal(gv, v) and np.array_equal(gf, want)


def _write_ply(path, v, faces, fmt, vertex_layout="xyz", strips=None, index_type="int", count_type="uchar"):
    """faces: list of index lists.  vertex_layout: 'xyz' | 'nxyz' (a float property in front and a uchar behind) | 'double'."""
    end = {"binary_little_endian": "<", "binary_big_endian": ">", "ascii": "<"}[fmt]
    vt = "double" if vertex_layout == "double" else "float"
    props = []
    if vertex_layout == "nxyz":
        props.append("property float confidence")
    props += ["property %s x" % vt, "property %s y" % vt, "property %s z" % vt]
    if vertex_layout == "nxyz":
        props.append("property uchar red")
    head = "ply\nformat %s 1.0\ncomment made by the tests\nelement vertex %d\n%s\n" % (fmt, len(v), "\n".join(props))
    if faces is not None:
        head += "element face %d\nproperty list %s %s vertex_indices\n" % (len(faces), count_type, index_type)
    if strips is not None:
        head += "element tristrips %d\nproperty list int int vertex_indices\n" % len(strips)
    head += "end_header\n"
    np_t = {"int": "i4", "uint": "u4", "short": "i2", "uchar": "u1", "int32": "i4"}
    with open(path, "wb") as fh:
        fh.write(head.encode())
        if fmt == "ascii":
            for p in v:
                row = ["%.9g" % c for c in p]
                if vertex_layout == "nxyz":
                    row = ["0.5"] + row + ["200"]
                fh.write((" ".join(row) + "\n").encode())
            for f in faces or []:
                fh.write(("%d %s\n" % (len(f), " ".join(map(str, f)))).encode())
            for s in strips or []:
                fh.write(("%d %s\n" % (len(s), " ".join(map(str, s)))).encode())
        else:
            for p in v:
                if vertex_layout == "nxyz":
                    fh.write(np.array([0.5], end + "f4").tobytes())
                fh.write(np.asarray(p, end + ("f8" if vt == "double" else "f4")).tobytes())
                if vertex_layout == "nxyz":
                    fh.write(b"\xc8")
            for f in faces or []:
                fh.write(np.array([len(f)], end + np_t[count_type]).tobytes() + np.asarray(f, end + np_t[index_type]).tobytes())
            for s in strips or []:
                fh.write(np.array([len(s)], end + "i4").tobytes() + np.asarray(s, end + "i4").tobytes())


def test_ply_variants_load_the_same_mesh(work, tmp_path):
    """ASCII / binary little- and big-endian PLY, extra vertex properties around x y z, double coordinates, other index types,
    files big enough to be parsed by several threads, at several thread counts: all give the bunny's arrays."""
    bv, bf = cases.mesh("bunny")
    reps = 12                                                   # > 1 MB of ASCII: the parallel paths really run
    v = np.concatenate([bv + np.float32(10 * k) for k in range(reps)])
    f = np.concatenate([bf + len(bv) * k for k in range(reps)]).astype(np.int32)
    faces = [list(t) for t in f]
    n = 0
    for fmt in ("ascii", "binary_little_endian", "binary_big_endian"):
        for layout in ("xyz", "nxyz", "double"):
            path = str(tmp_path / ("m%d.ply" % n))
            n += 1
            _write_ply(path, v, faces, fmt, layout)
            for threads in ("1", "5"):
                gv, gf = _dump(path, str(tmp_path / "p.bin"), {"VOXCLI_LOADER_THREADS": threads})
                assert np.array_equal(gv, v) and np.array_equal(gf, f), (fmt, layout, threads)
    path = str(tmp_path / "short.ply")
    _write_ply(path, bv, [list(t) for t in bf], "binary_little_endian", index_type="short", count_type="int")
    gv, gf = _dump(path, str(tmp_path / "p.bin"))
    assert np.array_equal(gv, bv) and np.array_equal(gf, bf)


def test_ply_triangle_strips(tmp_path):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 2, 0], [1, 2, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5]], np.float32)
    strips = [[0, 1, 2, 3, 4, 5, -1, 6, 7, 8]]
    want = np.array([[0, 1, 2], [2, 1, 3], [2, 3, 4], [4, 3, 5], [6, 7, 8]], np.int32)     # every second triangle flipped back
    for fmt in ("ascii", "binary_little_endian"):
        path = str(tmp_path / ("s_%s.ply" % fmt))
        _write_ply(path, v, None, fmt, strips=strips)
        gv, gf = _dump(path, str(tmp_path / "p.bin"))
        assert np.array_equal(gv, v) and np.array_equal(gf, want), fmt


def test_truncated_ply_is_an_error(tmp_path):
    bv, bf = cases.mesh("bunny")
    path = str(tmp_path / "t.ply")
    _write_ply(path, bv, [list(t) for t in bf], "binary_little_endian")
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:-100])
    r = subprocess.run([CLI, "-f", path, "-s", "64", "--dump-mesh", str(tmp_path / "x.bin")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "ends inside element" in r.stdout + r.stderr


def test_off_stl_3ds_load_the_same_mesh(tmp_path):
    """The other mesh formats the reference reads through trimesh2 (its help text names .3ds): OFF (with a quad and a comment), STL
    binary and ASCII (three fresh vertices per facet), 3DS (two objects, indices offset by the vertices before them)."""
    import struct
    bv, bf = cases.mesh("bunny")
    # OFF
    off = tmp_path / "m.off"
    quad_v = np.array([[0, 0, 9], [4, 0, 9], [4, 1, 9], [0, 1, 9]], np.float32)
    v = np.concatenate([bv, quad_v])
    with open(off, "w") as fh:
        fh.write("OFF\n# a comment\n%d %d 0\n" % (len(v), len(bf) + 1))
        for p in v:
            fh.write("%.9g %.9g %.9g\n" % tuple(p))
        for t in bf:
            fh.write("3 %d %d %d\n" % tuple(t))
        n = len(bv)
        fh.write("4 %d %d %d %d 255 0 0\n" % (n, n + 1, n + 2, n + 3))
    gv, gf = _dump(str(off), str(tmp_path / "o.bin"))
    assert np.array_equal(gv, v) and np.array_equal(gf[:-2], bf)
    assert np.array_equal(gf[-2:], np.array(_tess_quad(v, [n, n + 1, n + 2, n + 3]), np.int32))
    # STL, binary and ASCII
    soup = bv[bf.reshape(-1)].reshape(-1, 3, 3)
    stl = tmp_path / "m.stl"
    with open(stl, "wb") as fh:
        fh.write(b"solid looks like ascii but is binary".ljust(80, b" "))
        fh.write(struct.pack("<I", len(soup)))
        for tri in soup:
            fh.write(np.zeros(3, "<f4").tobytes() + tri.astype("<f4").tobytes() + b"\0\0")
    gv, gf = _dump(str(stl), str(tmp_path / "s.bin"))
    assert np.array_equal(gv, soup.reshape(-1, 3)) and np.array_equal(gf, np.arange(3 * len(soup), dtype=np.int32).reshape(-1, 3))
    stla = tmp_path / "a.stl"
    with open(stla, "w") as fh:
        fh.write("solid bunny\n")
        for tri in soup[:500]:
            fh.write(" facet normal 0 0 0\n  outer loop\n")
            for p in tri:
                fh.write("   vertex %.9g %.9g %.9g\n" % tuple(p))
            fh.write("  endloop\n endfacet\n")
        fh.write("endsolid bunny\n")
    gv, gf = _dump(str(stla), str(tmp_path / "s.bin"))
    assert np.array_equal(gv, soup[:500].reshape(-1, 3)) and len(gf) == 500
    # 3DS: two objects
    def chunk(cid, payload):
        return struct.pack("<HI", cid, 6 + len(payload)) + payload
    def obj(name, verts, faces):
        vl = struct.pack("<H", len(verts)) + verts.astype("<f4").tobytes()
        fl = struct.pack("<H", len(faces)) + b"".join(struct.pack("<4H", a, b, c, 0) for a, b, c in faces)
        return chunk(0x4000, name + b"\0" + chunk(0x4100, chunk(0x4110, vl) + chunk(0x4160, b"\0" * 48) + chunk(0x4120, fl)))
    half = 1200
    used1 = bf[(bf < half).all(axis=1)]
    v2, f2 = bv[half:], bf[(bf >= half).all(axis=1)] - half
    data = chunk(0x4D4D, chunk(0x0002, struct.pack("<I", 3)) + chunk(0x3D3D, obj(b"first", bv[:half], used1) + obj(b"second", v2, f2)))
    tds = tmp_path / "m.3ds"
    open(tds, "wb").write(data)
    gv, gf = _dump(str(tds), str(tmp_path / "t.bin"))
    assert np.array_equal(gv, bv)
    assert np.array_equal(gf, np.concatenate([used1, f2 + half]).astype(np.int32))
