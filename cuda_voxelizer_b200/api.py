"""Host-side mirror of the reference's voxelization interface, over the C ABI.

Names follow the reference: ``voxelize`` / ``voxelize_solid`` (main.cpp:23-24), ``voxinfo``
(util.h:50-69) -> :class:`Grid`, ``meshToGPU_managed`` (main.cpp:61-80) -> :func:`upload_soup` /
:func:`upload_indexed`.  PyTorch is used only as the owner of device memory and streams.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ACCUMULATE, MORTON, SOLID, TRIS_SOA4, Grid, Region, VoxError, check  # noqa: F401


def _f3(a):
    return (C.c_float * 3)(*[float(x) for x in a])


def device_count():
    n = C.c_int(0)
    check(_lib.lib().voxb200_device_count(C.byref(n)))
    return n.value


def init(device=0):
    check(_lib.lib().voxb200_init(int(device)))


def table_bytes(gridsize):
    return int(_lib.lib().voxb200_table_bytes(int(gridsize)))


def morton_encode(x, y, z):
    return int(_lib.lib().voxb200_morton_encode(int(x), int(y), int(z)))


def make_grid(mesh_min, mesh_max, gridsize, n_triangles):
    """createMeshBBCube + voxinfo ctor (util.h:56-61,80-110) on the host, bit-identical."""
    g = Grid()
    check(_lib.lib().voxb200_make_grid(_f3(mesh_min), _f3(mesh_max), int(gridsize), int(n_triangles), C.byref(g)))
    return g


def grid_from_verts(verts, gridsize, n_triangles):
    verts = np.asarray(verts, np.float32).reshape(-1, 3)
    return make_grid(verts.min(axis=0), verts.max(axis=0), gridsize, n_triangles)


def partition(gridsize, morton, part, n_parts):
    """Region of rank ``part``: z-slab (linear) or aligned morton block; returns (Region, bytes)."""
    r = Region()
    nbytes = C.c_size_t(0)
    check(_lib.lib().voxb200_partition(int(gridsize), int(bool(morton)), int(part), int(n_parts), C.byref(r), C.byref(nbytes)))
    return r, nbytes.value


def _stream_ptr(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))


def _run(fn, grid, tris, table, flags, region, stream):
    rp = C.byref(region) if region is not None else None
    check(fn(C.byref(grid), C.c_void_p(tris.data_ptr()), C.c_void_p(table.data_ptr()), int(flags), rp, _stream_ptr(stream)))
    return table


def _new_table(grid, region, morton, device):
    import torch
    if region is None:
        nbytes = table_bytes(grid.gridsize[0])
    else:
        sx, sy, sz = (region.hi[k] - region.lo[k] for k in range(3))
        nbytes = (sx * sy * sz + 31) // 32 * 4
    return torch.empty(nbytes // 4, dtype=torch.int32, device=device)


def voxelize(grid, tris, table=None, morton=False, region=None, stream=None, accumulate=False, soa4=False):
    """Surface voxelization (reference: voxelize(), voxelize.cu:192).  ``tris``: CUDA float32 tensor,
    9 floats per triangle (or 3 float4 planes with ``soa4``).  Returns the uint32 bit table as an
    int32 CUDA tensor (the region's words only when ``region`` is given).  Asynchronous on ``stream``."""
    if table is None:
        table = _new_table(grid, region, morton, getattr(tris, "device", "cuda"))
    flags = (MORTON if morton else 0) | (ACCUMULATE if accumulate else 0) | (TRIS_SOA4 if soa4 else 0)
    return _run(_lib.lib().voxb200_surface, grid, tris, table, flags, region, stream)


def voxelize_solid(grid, tris, table=None, morton=False, region=None, stream=None, accumulate=False, soa4=False):
    """Solid voxelization (reference: voxelize_solid(), voxelize_solid.cu:147)."""
    if table is None:
        table = _new_table(grid, region, morton, getattr(tris, "device", "cuda"))
    flags = (MORTON if morton else 0) | (ACCUMULATE if accumulate else 0) | (TRIS_SOA4 if soa4 else 0)
    return _run(_lib.lib().voxb200_solid, grid, tris, table, flags, region, stream)


def voxelize_host(grid, host_tris, host_table=None, solid=False, morton=False, region=None):
    """End to end with host buffers (numpy or pinned torch CPU tensors): H2D + voxelize + D2H.
    Returns (table, timing_ms[h2d, voxelize, d2h, total])."""
    def ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
    if host_table is None:
        if region is None:
            nbytes = table_bytes(grid.gridsize[0])
        else:
            sx, sy, sz = (region.hi[k] - region.lo[k] for k in range(3))
            nbytes = (sx * sy * sz + 31) // 32 * 4
        host_table = np.empty(nbytes // 4, np.uint32)
    timing = (C.c_float * 4)()
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    rp = C.byref(region) if region is not None else None
    check(_lib.lib().voxb200_voxelize_host(C.byref(grid), C.c_void_p(ptr(host_tris)), C.c_void_p(ptr(host_table)), flags, rp, timing))
    return host_table, [float(t) for t in timing]


class DeviceBuffer:
    """A device allocation made by the library (voxb200_upload_*); freed on close()/GC."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def data_ptr(self):
        return self.ptr

    def close(self):
        if self.ptr:
            try:
                _lib.lib().voxb200_free(C.c_void_p(self.ptr))
            except TypeError:          # interpreter shutdown: the module globals are gone, the process is about to release the memory
                pass
            self.ptr = 0

    __del__ = close


def upload_soup(host_tris, soa4=False, stream=0):
    """Triangle soup upload (reference: meshToGPU_managed, main.cpp:61-80)."""
    a = np.ascontiguousarray(host_tris, np.float32).reshape(-1, 9)
    out = C.c_void_p(0)
    check(_lib.lib().voxb200_upload_soup(C.c_void_p(a.ctypes.data), len(a), int(soa4), C.byref(out), C.c_void_p(stream)))
    return DeviceBuffer(out.value, len(a) * (48 if soa4 else 36))


def upload_indexed(verts, faces, soa4=False, want_bbox=True, stream=0):
    """Indexed mesh upload with device-side expansion and bbox reduction.  Returns (buffer, min, max)."""
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
    out = C.c_void_p(0)
    mn, mx = (C.c_float * 3)(), (C.c_float * 3)()
    check(_lib.lib().voxb200_upload_indexed(C.c_void_p(v.ctypes.data), len(v), C.c_void_p(f.ctypes.data), len(f), int(soa4),
                                            C.byref(out), mn if want_bbox else None, mx if want_bbox else None, C.c_void_p(stream)))
    return DeviceBuffer(out.value, len(f) * (48 if soa4 else 36)), np.array(mn[:], np.float32), np.array(mx[:], np.float32)


def download(buf_ptr, nbytes, stream=0):
    out = np.empty(nbytes // 4, np.uint32)
    check(_lib.lib().voxb200_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(buf_ptr), nbytes, C.c_void_p(stream)))
    return out


def download_table(d_table, host_table, stream=None):
    """voxb200_download_table: CUDA tensor -> host array (numpy, or a pinned torch CPU tensor for full speed); sparse read-back
    when it pays.  Returns (host_table, {"sparse": bool, "nonzero_words": int})."""
    info = (C.c_uint64 * 2)()
    hp = host_table.data_ptr() if hasattr(host_table, "data_ptr") else host_table.ctypes.data
    check(_lib.lib().voxb200_download_table(C.c_void_p(d_table.data_ptr()), d_table.numel(), C.c_void_p(hp), _stream_ptr(stream), info))
    return host_table, {"sparse": bool(info[0]), "nonzero_words": int(info[1])}


def set_readback_mode(mode):
    """'auto' | 'dense' | 'sparse' (voxb200_set_readback_mode)."""
    check(_lib.lib().voxb200_set_readback_mode({"auto": 0, "dense": 1, "sparse": 2}[mode]))


def binvox_rle(d_table, gridsize, stream=None):
    """voxb200_binvox_rle: the binvox payload (numpy uint8) of a linear CUDA table, run-length encoded on the device."""
    ptr, n = C.c_void_p(), C.c_size_t()
    check(_lib.lib().voxb200_binvox_rle(C.c_void_p(d_table.data_ptr()), int(gridsize), C.byref(ptr), C.byref(n), _stream_ptr(stream)))
    out = np.empty(n.value, np.uint8)
    try:
        if n.value:
            check(_lib.lib().voxb200_memcpy_d2h(C.c_void_p(out.ctypes.data), ptr, n.value, None))
    finally:
        _lib.lib().voxb200_free(ptr)
    return out


def set_host_threads(n):
    check(_lib.lib().voxb200_set_host_threads(int(n)))


def last_readback():
    info = (C.c_uint64 * 2)()
    check(_lib.lib().voxb200_last_readback(info))
    return {"sparse": bool(info[0]), "nonzero_words": int(info[1])}


def launch_count(reset=False):
    return int(_lib.lib().voxb200_launch_count(int(reset)))


def last_counters():
    out = (C.c_uint64 * 4)()
    check(_lib.lib().voxb200_last_counters(out))
    return {"coop_triangles": int(out[0]), "coop_items": int(out[1]), "solid_clamped": int(out[2]), "solid_row_lists": int(out[3])}


def set_profiling(on):
    check(_lib.lib().voxb200_set_profiling(int(bool(on))))


def phase_ms(call_index):
    """[zero-fill, per-triangle kernel, cooperative kernel, solid scan] ms of the i-th profiled call."""
    out = (C.c_float * 4)()
    check(_lib.lib().voxb200_phase_ms(int(call_index), out))
    return [float(x) for x in out]


def route_triangles(grid, tris, region, solid=False, morton=False, stream=0):
    """Device-side routing of a 9-float soup to one region (multi-GPU slabs).  Returns (DeviceBuffer, count)."""
    out = C.c_void_p(0)
    n = C.c_size_t(0)
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    check(_lib.lib().voxb200_route_triangles(C.byref(grid), C.c_void_p(tris.data_ptr()), flags, C.byref(region), C.byref(out), C.byref(n), C.c_void_p(stream)))
    return DeviceBuffer(out.value, n.value * 36), int(n.value)


def sort_triangles(grid, tris, stream=None):
    """Upload-path option: a copy of the device soup (CUDA float32 tensor / DeviceBuffer, 9 floats per triangle) ordered by
    the z-layer of each triangle's lowest vertex.  Returns a DeviceBuffer."""
    out = C.c_void_p(0)
    check(_lib.lib().voxb200_sort_triangles(C.byref(grid), C.c_void_p(tris.data_ptr()), C.byref(out), _stream_ptr(stream)))
    return DeviceBuffer(out.value, int(grid.n_triangles) * 36)


def voxelize_host_indexed(grid, host_verts, host_faces, host_table=None, solid=False, morton=False, region=None):
    """End to end from the indexed mesh (numpy or pinned torch CPU tensors): H2D of vertices + faces, expansion on the
    GPU, voxelization, D2H.  Returns (table, timing_ms[h2d+expand, voxelize, d2h, total])."""
    def ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
    n_verts = (host_verts.numel() if hasattr(host_verts, "numel") else host_verts.size) // 3
    if host_table is None:
        if region is None:
            nbytes = table_bytes(grid.gridsize[0])
        else:
            sx, sy, sz = (region.hi[k] - region.lo[k] for k in range(3))
            nbytes = (sx * sy * sz + 31) // 32 * 4
        host_table = np.empty(nbytes // 4, np.uint32)
    timing = (C.c_float * 4)()
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    rp = C.byref(region) if region is not None else None
    check(_lib.lib().voxb200_voxelize_host_indexed(C.byref(grid), C.c_void_p(ptr(host_verts)), n_verts, C.c_void_p(ptr(host_faces)),
                                                   C.c_void_p(ptr(host_table)), flags, rp, timing))
    return host_table, [float(t) for t in timing]


def voxelize_host_nonzero(grid, host_verts, host_faces, solid=False, morton=False, region=None):
    """voxb200_voxelize_host_nonzero: end to end from the indexed mesh to the table's non-zero words.  Returns (pairs, timing_ms):
    pairs = uint32 array [n, 2] of {word index, bits}, ascending (a COPY of the library's pinned buffer)."""
    def ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
    n_verts = (host_verts.numel() if hasattr(host_verts, "numel") else host_verts.size) // 3
    out, n = C.c_void_p(), C.c_size_t()
    timing = (C.c_float * 4)()
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    rp = C.byref(region) if region is not None else None
    check(_lib.lib().voxb200_voxelize_host_nonzero(C.byref(grid), C.c_void_p(ptr(host_verts)), n_verts, C.c_void_p(ptr(host_faces)), flags, rp,
                                                   C.byref(out), C.byref(n), timing))
    pairs = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(n.value, 2)).copy() if n.value else np.zeros((0, 2), np.uint32)
    return pairs, [float(t) for t in timing]


def route_triangles_multi(grid, tris, regions, out, solid=False, morton=False, stream=None):
    """Route a device soup to several regions at once into the preallocated CUDA tensor ``out`` (float32, capacity
    out.numel() // 9 triangles), segments back to back in region order.  Returns the per-region triangle counts."""
    arr = (Region * len(regions))(*regions)
    counts = (C.c_size_t * len(regions))()
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    check(_lib.lib().voxb200_route_triangles_multi(C.byref(grid), C.c_void_p(tris.data_ptr()), flags, arr, len(regions),
                                                   C.c_void_p(out.data_ptr()), out.numel() // 9, counts, _stream_ptr(stream)))
    return [int(c) for c in counts]


def extract_voxels(table, first_voxel=0, stream=None):
    """Device-side compaction of a table (CUDA int32 tensor) into the ascending indices of its set voxels.
    Returns a numpy uint64 array (copied to the host)."""
    out = C.c_void_p(0)
    n = C.c_size_t(0)
    check(_lib.lib().voxb200_extract_voxels(C.c_void_p(table.data_ptr()), table.numel(), int(first_voxel), C.byref(out), C.byref(n), _stream_ptr(stream)))
    buf = DeviceBuffer(out.value, n.value * 8)
    host = np.empty(n.value, np.uint64)
    if n.value:
        check(_lib.lib().voxb200_memcpy_d2h(C.c_void_p(host.ctypes.data), C.c_void_p(buf.ptr), n.value * 8, _stream_ptr(stream)))
    buf.close()
    return host


def release():
    """Free everything the library cached for the current device."""
    check(_lib.lib().voxb200_release())


class Mesh:
    """A mesh prepared for repeated voxelization on one grid (voxb200_mesh_*): owns its re-ordered triangles, the tile plan
    and a private workspace.  ``tris``: CUDA float32 tensor with 9 floats per triangle, or pass ``verts`` / ``faces`` (CUDA
    float32 [V,3] / int32 [T,3]) for an indexed mesh.  Re-entrant: meshes may voxelize concurrently on different streams."""

    def __init__(self, grid, tris=None, verts=None, faces=None, solid=False, morton=False, region=None, stream=None):
        self.grid, self.region, self.solid, self.morton = grid, region, solid, morton
        self._h = C.c_void_p(0)
        flags = (MORTON if morton else 0) | (SOLID if solid else 0)
        rp = C.byref(region) if region is not None else None
        if tris is not None:
            check(_lib.lib().voxb200_mesh_create(C.byref(grid), C.c_void_p(tris.data_ptr()), flags, rp, C.byref(self._h), _stream_ptr(stream)))
        else:
            check(_lib.lib().voxb200_mesh_create_indexed(C.byref(grid), C.c_void_p(verts.data_ptr()), verts.numel() // 3, C.c_void_p(faces.data_ptr()),
                                                         flags, rp, C.byref(self._h), _stream_ptr(stream)))

    def update(self, tris=None, verts=None, faces=None, stream=None):
        """New vertex positions (same triangle count): re-prepares the handle in place."""
        if tris is not None:
            check(_lib.lib().voxb200_mesh_update(self._h, C.c_void_p(tris.data_ptr()), _stream_ptr(stream)))
        else:
            check(_lib.lib().voxb200_mesh_update_indexed(self._h, C.c_void_p(verts.data_ptr()), verts.numel() // 3, C.c_void_p(faces.data_ptr()), _stream_ptr(stream)))

    def voxelize(self, table=None, accumulate=False, stream=None):
        if table is None:
            table = _new_table(self.grid, self.region, self.morton, "cuda")
        check(_lib.lib().voxb200_mesh_voxelize(self._h, C.c_void_p(table.data_ptr()), ACCUMULATE if accumulate else 0, _stream_ptr(stream)))
        return table

    def info(self):
        out = (C.c_uint64 * 8)()
        check(_lib.lib().voxb200_mesh_info(self._h, out))
        keys = ("tile_schedule", "tiles", "work_tiles", "instances", "side_triangles", "batches", "zero_quota", "zero_blocks")
        return dict(zip(keys, (int(x) for x in out)))

    def counters(self):
        out = (C.c_uint64 * 4)()
        check(_lib.lib().voxb200_mesh_counters(self._h, out))
        return {"coop_triangles": int(out[0]), "coop_items": int(out[1]), "solid_clamped": int(out[2]), "solid_row_lists": int(out[3])}

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and C is not None:          # (module globals are gone at interpreter exit)
            _lib.lib().voxb200_mesh_destroy(h)

    __del__ = close


def voxelize_host_multi(grid, host_verts, host_faces, host_table=None, solid=False, morton=False, devices=None, n_devices=None):
    """One process, N GPUs (voxb200_voxelize_host_multi): indexed mesh and table in host memory (numpy, or pinned torch CPU
    tensors for full speed).  Returns (table, timing_ms[h2d, peer gather, prepare, voxelize, d2h, total, wall, n])."""
    def ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
    n_verts = (host_verts.numel() if hasattr(host_verts, "numel") else host_verts.size) // 3
    if devices is None:
        devices = list(range(n_devices if n_devices is not None else device_count()))
    if host_table is None:
        host_table = np.empty(table_bytes(grid.gridsize[0]) // 4, np.uint32)
    arr = (C.c_int * len(devices))(*devices)
    timing = (C.c_float * 8)()
    flags = (MORTON if morton else 0) | (SOLID if solid else 0)
    check(_lib.lib().voxb200_voxelize_host_multi(C.byref(grid), C.c_void_p(ptr(host_verts)), n_verts, C.c_void_p(ptr(host_faces)),
                                                 C.c_void_p(ptr(host_table)), flags, arr, len(devices), timing))
    return host_table, [float(t) for t in timing]


def gather_slabs(slabs, devices, out, out_device, stream=None):
    """voxb200_gather_slabs: device-resident slabs (CUDA tensors, region order) -> one CUDA tensor `out` on `out_device`."""
    n = len(slabs)
    ptrs = (C.c_void_p * n)(*[s.data_ptr() for s in slabs])
    devs = (C.c_int * n)(*devices)
    sizes = (C.c_size_t * n)(*[s.numel() * s.element_size() for s in slabs])
    check(_lib.lib().voxb200_gather_slabs(ptrs, devs, sizes, n, C.c_void_p(out.data_ptr()), int(out_device), _stream_ptr(stream)))
    return out
