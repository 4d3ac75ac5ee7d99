"""cuda_voxelizer_b200 — B200-native (sm_100a) voxelization hot path behind the reference's API.

The product is the shared library ``libvoxb200.so`` (C ABI: include/voxb200.h; C++ drop-in symbols:
include/voxelize_dropin.h) and the ``bin/cuda_voxelizer`` CLI.  This package is the thin Python host
mirror used by the tests and the benchmark.
"""
from . import meshgen, meshio  # noqa: F401
from ._lib import ACCUMULATE, MORTON, SOLID, TRIS_SOA4, Grid, Region, VoxError  # noqa: F401
from .api import (Mesh, binvox_rle, device_count, download_table, gather_slabs, voxelize_host_multi, download, extract_voxels, grid_from_verts, init, last_counters, last_readback, set_readback_mode, set_host_threads, launch_count, make_grid,  # noqa: F401
                  morton_encode, partition, phase_ms, release, route_triangles, route_triangles_multi, set_profiling, sort_triangles, table_bytes, upload_indexed, upload_soup, voxelize, voxelize_host, voxelize_host_indexed, voxelize_host_nonzero,
                  voxelize_solid)
