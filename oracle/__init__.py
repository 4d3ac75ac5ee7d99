"""TEST INFRASTRUCTURE — not product code.

ctypes access to the two CPU checkers:

* ``liboracle.so``       — oracle/vox_oracle.c, the plain-C restatement of the reference algorithm
* ``_ref/libvoxref.so``  — the reference's own unmodified ``src/cpu_voxelizer.cpp`` compiled here
                           from /root/reference (oracle/Makefile); travels to the GPU box prebuilt
* ``_ref/libvoxref_gpu.so`` — the reference's own unmodified GPU kernels (``src/voxelize.cu``,
                           ``src/voxelize_solid.cu``) compiled for sm_100a: bench.py's second baseline

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``cuda_voxelizer_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libvoxref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (always possible) and _ref/libvoxref.so (when /root/reference exists)."""
    need = force or not os.path.exists(_ORACLE_SO) or (
        os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "vox_oracle.c")))
    need_ref = os.path.exists("/root/reference/src/cpu_voxelizer.cpp") and (
        force or not os.path.exists(_REF_SO)
        or os.path.getmtime(_REF_SO) < os.path.getmtime(os.path.join(_HERE, "ref_driver.cpp")))
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    if need_ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.oracle_bbox_cube.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.oracle_mesh_bbox.argtypes = [_f32p, C.c_size_t, _f32p, _f32p]
        L.oracle_unit.argtypes = [_f32p, _f32p, C.c_uint, _f32p]
        L.oracle_table_bytes.argtypes = [C.c_uint]
        L.oracle_table_bytes.restype = C.c_size_t
        L.oracle_reference_table_bytes.argtypes = [C.c_uint]
        L.oracle_reference_table_bytes.restype = C.c_size_t
        L.oracle_morton.argtypes = [C.c_uint, C.c_uint, C.c_uint]
        L.oracle_morton.restype = C.c_uint64
        for fn in (L.oracle_surface, L.oracle_solid):
            fn.argtypes = [_f32p, C.c_size_t, _f32p, _f32p, C.c_uint, C.c_int, C.c_int, C.c_int, _u32p, C.c_void_p]
            fn.restype = None
        L.oracle_solid_ub_events.restype = C.c_uint64
        L.oracle_set_table_origin.argtypes = [C.c_size_t]
        L.oracle_set_table_origin.restype = None
        L.oracle_expand_soup.argtypes = [_f32p, _i32p, C.c_size_t, _f32p]
        L.oracle_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_fnv1a64.restype = C.c_uint64
        L.oracle_popcount.argtypes = [_u32p, C.c_size_t]
        L.oracle_popcount.restype = C.c_uint64
        L.oracle_check_voxel.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, _u32p]
        L.oracle_check_voxel.restype = C.c_int
        _lib = L
    return _lib


def have_ref():
    if not os.path.exists(_REF_SO):
        try:
            build()
        except Exception:
            pass
    return os.path.exists(_REF_SO)


def ref():
    """The compiled, unmodified reference CPU voxelizer (None-safe: raises if not built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libvoxref.so missing (built only where /root/reference exists)")
        R = C.CDLL(_REF_SO)
        R.voxref_voxinfo.argtypes = [_f32p, C.c_size_t, C.c_uint, C.c_size_t, _f32p]
        R.voxref_voxinfo.restype = C.c_int
        R.voxref_voxinfo_layout.argtypes = [_u64p]
        R.voxref_voxelize.argtypes = [_f32p, C.c_size_t, _i32p, C.c_size_t, C.c_uint, C.c_int, C.c_int, _u32p, C.c_int]
        R.voxref_voxelize.restype = C.c_double
        R.voxref_max_threads.restype = C.c_int
        R.voxref_set_threads.argtypes = [C.c_int]
        _ref = R
    return _ref


# ----------------------------------------------------------------------------- restatement oracle

def table_words(gridsize):
    return lib().oracle_table_bytes(gridsize) // 4


def voxinfo(verts, gridsize):
    """(bbox_min, bbox_max, unit) exactly as main.cpp:179-186 + util.h:56-61,80-110 derive them."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().oracle_mesh_bbox(verts, len(verts), mn, mx)
    cmn, cmx = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().oracle_bbox_cube(mn, mx, cmn, cmx)
    unit = np.zeros(3, np.float32)
    lib().oracle_unit(cmn, cmx, gridsize, unit)
    return cmn, cmx, unit


def soup(verts, faces):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    out = np.empty((len(faces), 9), np.float32)
    lib().oracle_expand_soup(verts, faces, len(faces), out)
    return out


def _run(fn, tris, bb_min, unit, gridsize, morton, z_range, table):
    tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
    if table is None:
        table = np.zeros(table_words(gridsize), np.uint32)
    z0, z1 = (0, gridsize) if z_range is None else z_range
    stats = np.zeros(2, np.uint64)
    fn(tris, len(tris), np.ascontiguousarray(bb_min, np.float32), np.ascontiguousarray(unit, np.float32),
       gridsize, int(bool(morton)), int(z0), int(z1), table, stats.ctypes.data)
    return table, stats


def surface(tris, bb_min, unit, gridsize, morton=False, z_range=None, table=None, return_stats=False):
    t, s = _run(lib().oracle_surface, tris, bb_min, unit, gridsize, morton, z_range, table)
    return (t, s) if return_stats else t


def solid(tris, bb_min, unit, gridsize, morton=False, z_range=None, table=None, return_stats=False):
    t, s = _run(lib().oracle_solid, tris, bb_min, unit, gridsize, morton, z_range, table)
    return (t, s) if return_stats else t


def surface_slab(tris, bb_min, unit, gridsize, z0, z1):
    """The words of z-slab [z0, z1) of the LINEAR surface table of a grid that may be too big for host memory as a whole
    (8192^3 = 64 GiB): the oracle run clipped to the slab, writing into a slab-sized table."""
    words_per_layer = gridsize * gridsize // 32
    table = np.zeros(words_per_layer * (z1 - z0), np.uint32)
    lib().oracle_set_table_origin(words_per_layer * z0)
    try:
        surface(tris, bb_min, unit, gridsize, z_range=(z0, z1), table=table)
    finally:
        lib().oracle_set_table_origin(0)
    return table


def solid_ub_events():
    return int(lib().oracle_solid_ub_events())


def morton(x, y, z):
    return int(lib().oracle_morton(int(x), int(y), int(z)))


def fnv1a64(arr):
    arr = np.ascontiguousarray(arr)
    return int(lib().oracle_fnv1a64(arr.ctypes.data, arr.nbytes))


def popcount(table):
    table = np.ascontiguousarray(table, dtype=np.uint32)
    return int(lib().oracle_popcount(table, table.size))


# ----------------------------------------------------------------------------- compiled reference

def ref_voxinfo(verts, gridsize, n_tris=0):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    out = np.zeros(16, np.float32)
    ref().voxref_voxinfo(verts, len(verts), gridsize, n_tris, out)
    return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()


def ref_voxinfo_layout():
    out = np.zeros(6, np.uint64)
    ref().voxref_voxinfo_layout(out)
    return dict(zip(("sizeof", "bbox", "gridsize", "n_triangles", "unit", "alignof"), (int(v) for v in out)))


def ref_voxelize(verts, faces, gridsize, solid=False, morton=False, threads=None, return_ms=False):
    """Run the reference's own cpu_voxelize_mesh{,_solid} on an indexed mesh; returns the table."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    table = np.zeros(table_words(gridsize), np.uint32)
    R = ref()
    if threads is not None:
        R.voxref_set_threads(int(threads))
    ms = R.voxref_voxelize(verts, len(verts), faces, len(faces), gridsize, int(bool(solid)), int(bool(morton)), table, 1)
    return (table, ms) if return_ms else table


def ref_max_threads():
    return int(ref().voxref_max_threads())


# ----------------------------------------------------------------------------- compiled reference writers
_REF_IO_SO = os.path.join(_HERE, "_ref", "libvoxref_io.so")
_ref_io = None


def have_ref_io():
    return os.path.exists(_REF_IO_SO)


def ref_write(fmt, table, gridsize, bbox_min, bbox_max, n_tris, base_filename):
    """Run one of the reference's own writers (util_io.cpp): fmt in binvox|morton|obj_points|obj|vox."""
    global _ref_io
    if _ref_io is None:
        _ref_io = C.CDLL(_REF_IO_SO)
        _ref_io.voxref_write.argtypes = [C.c_int, _u32p, C.c_uint, _f32p, _f32p, C.c_size_t, C.c_char_p]
    code = {"binvox": 0, "morton": 1, "obj_points": 2, "obj": 3, "vox": 4}[fmt]
    rc = _ref_io.voxref_write(code, np.ascontiguousarray(table, np.uint32), gridsize, np.ascontiguousarray(bbox_min, np.float32),
                              np.ascontiguousarray(bbox_max, np.float32), n_tris, base_filename.encode())
    assert rc == 0


# ----------------------------------------------------------------------------- compiled reference GPU kernels (bench baseline)
_REF_GPU_SO = os.path.join(_HERE, "_ref", "libvoxref_gpu.so")
_ref_gpu = None


def have_ref_gpu():
    return os.path.exists(_REF_GPU_SO)


def ref_gpu_run(bbox_min, bbox_max, gridsize, soup9, solid=False, morton=False, warmup=2, reps=5, want_table=False):
    """Time the reference's own voxelize() / voxelize_solid() (unmodified kernels, sm_100a build) on the current CUDA
    device with triangles and table device-resident.  Returns ({"mean_ms", "best_ms", "memset_ms"}, table or None)."""
    global _ref_gpu
    if _ref_gpu is None:
        _ref_gpu = C.CDLL(_REF_GPU_SO)
        _ref_gpu.voxrefgpu_run.argtypes = [_f32p, C.c_uint, _f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p]
    soup9 = np.ascontiguousarray(soup9, np.float32).reshape(-1, 9)
    bbox6 = np.ascontiguousarray(np.concatenate([np.asarray(bbox_min, np.float32), np.asarray(bbox_max, np.float32)]))
    out = np.zeros(3, np.float32)
    table = np.zeros(table_words(gridsize), np.uint32) if want_table else None
    rc = _ref_gpu.voxrefgpu_run(bbox6, gridsize, soup9, len(soup9), int(bool(solid)), int(bool(morton)), int(warmup), int(reps), out,
                                table.ctypes.data if table is not None else None)
    if rc != 0:
        raise RuntimeError("reference GPU kernels failed: CUDA error %d" % rc)
    return {"mean_ms": float(out[0]), "best_ms": float(out[1]), "memset_ms": float(out[2])}, table
