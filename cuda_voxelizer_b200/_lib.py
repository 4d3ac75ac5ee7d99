"""ctypes binding of libvoxb200.so (the C ABI in include/voxb200.h).

There is no Python or CPU fallback: if the shared library is missing this module raises, and every
compute call fails with VOXB200_ENODEVICE when no sm_100 GPU is present.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvoxb200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ENODEVICE, ECUDA, EINVAL, ENOMEM = 0, 1, 2, 3, 4
MORTON, ACCUMULATE, TRIS_SOA4, SOLID = 1, 2, 4, 8

# every symbol include/voxb200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "voxb200_device_count", "voxb200_init", "voxb200_last_error", "voxb200_make_grid", "voxb200_table_bytes",
    "voxb200_partition", "voxb200_morton_encode", "voxb200_malloc", "voxb200_free", "voxb200_memcpy_d2h",
    "voxb200_upload_soup", "voxb200_upload_indexed", "voxb200_surface", "voxb200_solid", "voxb200_voxelize_host",
    "voxb200_launch_count", "voxb200_last_counters", "voxb200_version", "voxb200_set_profiling", "voxb200_phase_ms",
    "voxb200_route_triangles", "voxb200_voxelize_host_indexed", "voxb200_route_triangles_multi", "voxb200_extract_voxels", "voxb200_release", "voxb200_sort_triangles",
    "voxb200_reference_table_bytes", "voxb200_mesh_create", "voxb200_mesh_create_indexed", "voxb200_mesh_update", "voxb200_mesh_update_indexed",
    "voxb200_mesh_voxelize", "voxb200_mesh_info", "voxb200_mesh_destroy", "voxb200_mesh_counters",
    "voxb200_voxelize_host_multi", "voxb200_gather_slabs", "voxb200_host_alloc", "voxb200_host_free", "voxb200_download_table", "voxb200_voxelize_host_nonzero", "voxb200_binvox_rle", "voxb200_last_readback", "voxb200_set_readback_mode", "voxb200_set_host_threads", "voxb200_selftest_host_pool",
]


class Grid(C.Structure):
    """voxb200_grid == the reference's voxinfo (util.h:50-69), 64 bytes."""
    _fields_ = [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("gridsize", C.c_uint * 3),
                ("n_triangles", C.c_size_t), ("unit", C.c_float * 3)]


class Region(C.Structure):
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3)]


class VoxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("voxb200 error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile libvoxb200.so (+ the CLI) in-tree with nvcc for sm_100a."""
    out = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libvoxb200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("libvoxb200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C cuda_voxelizer_b200/csrc` — there is no fallback path" % SO_PATH)
    L = C.CDLL(SO_PATH)
    f3 = C.POINTER(C.c_float)
    L.voxb200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.voxb200_init.argtypes = [C.c_int]
    L.voxb200_last_error.restype = C.c_char_p
    L.voxb200_version.restype = C.c_char_p
    L.voxb200_make_grid.argtypes = [f3, f3, C.c_uint, C.c_size_t, C.POINTER(Grid)]
    L.voxb200_table_bytes.argtypes = [C.c_uint]
    L.voxb200_table_bytes.restype = C.c_size_t
    L.voxb200_reference_table_bytes.argtypes = [C.c_uint]
    L.voxb200_reference_table_bytes.restype = C.c_size_t
    L.voxb200_partition.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.POINTER(Region), C.POINTER(C.c_size_t)]
    L.voxb200_morton_encode.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.voxb200_morton_encode.restype = C.c_uint64
    L.voxb200_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.voxb200_free.argtypes = [C.c_void_p]
    L.voxb200_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.voxb200_upload_soup.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]
    L.voxb200_upload_indexed.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), f3, f3, C.c_void_p]
    for fn in (L.voxb200_surface, L.voxb200_solid):
        fn.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(Region), C.c_void_p]
    L.voxb200_voxelize_host.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(Region), f3]
    L.voxb200_route_triangles.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_uint, C.POINTER(Region), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    L.voxb200_sort_triangles.argtypes = [C.POINTER(Grid), C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    L.voxb200_voxelize_host_indexed.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(Region), f3]
    L.voxb200_route_triangles_multi.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_uint, C.POINTER(Region), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
    L.voxb200_extract_voxels.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    L.voxb200_mesh_create.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_uint, C.POINTER(Region), C.POINTER(C.c_void_p), C.c_void_p]
    L.voxb200_mesh_create_indexed.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.POINTER(Region), C.POINTER(C.c_void_p), C.c_void_p]
    L.voxb200_mesh_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.voxb200_mesh_update_indexed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.voxb200_mesh_voxelize.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    L.voxb200_mesh_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.voxb200_mesh_destroy.argtypes = [C.c_void_p]
    L.voxb200_mesh_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.voxb200_voxelize_host_multi.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(C.c_int), C.c_int, f3]
    L.voxb200_gather_slabs.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.voxb200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.voxb200_host_free.argtypes = [C.c_void_p]
    L.voxb200_set_readback_mode.argtypes = [C.c_int]
    L.voxb200_set_host_threads.argtypes = [C.c_int]
    L.voxb200_binvox_rle.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    L.voxb200_voxelize_host_nonzero.argtypes = [C.POINTER(Grid), C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), f3]
    L.voxb200_last_readback.argtypes = [C.POINTER(C.c_uint64)]
    L.voxb200_download_table.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.voxb200_launch_count.argtypes = [C.c_int]
    L.voxb200_launch_count.restype = C.c_uint64
    L.voxb200_last_counters.argtypes = [C.POINTER(C.c_uint64)]
    L.voxb200_set_profiling.argtypes = [C.c_int]
    L.voxb200_phase_ms.argtypes = [C.c_uint, f3]
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise VoxError(rc, lib().voxb200_last_error().decode("utf-8", "replace"))
