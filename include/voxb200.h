/*
 * voxb200.h — C ABI of the B200-native voxelization hot path.
 *
 * This is the drop-in boundary for the hot path of Forceflow/cuda_voxelizer.
 * Plain pointers and sizes only; every function returns 0 on success or a non-zero VOXB200_E*
 * code, with a message retrievable through voxb200_last_error().  The reference reports CUDA
 * failures by printing and exit(EXIT_FAILURE) (src/libs/cuda/helper_cuda.h:566-579); the C++
 * drop-in symbols in voxelize_dropin.h reproduce that on top of these status codes.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device every compute entry point fails
 * with VOXB200_ENODEVICE.
 *
 * Threading: like the reference (which writes __constant__ LUTs and uses the legacy default stream), the library
 * keeps per-device scratch (work queue, counters, staging buffers) and is NOT re-entrant: use one host thread per
 * device, or serialise calls that target the same device.  Calls on different streams of one device must not
 * overlap for the same reason.
 *
 * Reference interfaces replaced (file:line into the reference tree):
 *   voxb200_init / voxb200_device_count ... initCuda()                    src/util_cuda.cpp:4-42
 *   voxb200_make_grid .................... createMeshBBCube + voxinfo     src/util.h:56-61, 80-110 (main.cpp:184-186)
 *   voxb200_table_bytes .................. vtable_size                    src/main.cpp:190
 *   voxb200_upload_soup / _indexed ....... meshToGPU_managed()            src/main.cpp:61-80
 *   voxb200_sort_triangles ............... (new) optional z-layer ordering of the uploaded soup
 *   voxb200_surface ...................... voxelize()                     src/voxelize.cu:192-238 (kernel :58-190)
 *   voxb200_solid ........................ voxelize_solid()               src/voxelize_solid.cu:147-193 (kernel :73-145)
 *   voxb200_morton_encode ................ mortonEncode_LUT()             src/voxelize.cuh:20-34
 *   voxb200_voxelize_host ................ main.cpp:203-222 (upload + voxelize + table read-back by the writers)
 *   voxb200_voxelize_host_multi .......... main.cpp:203-222 with N GPUs (new: the reference is single-GPU)
 *   voxb200_mesh_* ....................... (new) prepared mesh for repeated voxelization (README.md:74 "per-frame")
 *   bit-table layout ..................... setBit / checkVoxel            src/voxelize.cu:50-55, src/util.h:25-38
 */
#ifndef VOXB200_H
#define VOXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define VOXB200_OK            0
#define VOXB200_ENODEVICE     1   /* no CUDA device / device is not sm_100                          */
#define VOXB200_ECUDA         2   /* a CUDA runtime call failed (see voxb200_last_error)            */
#define VOXB200_EINVAL        3   /* bad argument (NULL pointer, empty region, unsupported grid)    */
#define VOXB200_ENOMEM        4   /* device or pinned-host allocation failed                        */

/* ---- flags for voxb200_surface / voxb200_solid / voxb200_voxelize_host --------------------- */
#define VOXB200_MORTON        1u  /* morton-ordered table (reference: -o morton, morton_code=true)  */
#define VOXB200_ACCUMULATE    2u  /* OR / XOR into the table's current content instead of clearing  */
                                  /* it first — the reference's exact semantics (it never clears,   */
                                  /* main.cpp:214)                                                  */
#define VOXB200_TRIS_SOA4     4u  /* triangles are 3 planes of float4 (v0|v1|v2, .w ignored), each  */
                                  /* plane n_triangles long — the layout voxb200_upload_* produce   */
                                  /* with soa4=1.  Default: the reference's 9-float AoS records.    */
#define VOXB200_SOLID         8u  /* voxb200_voxelize_host only: solid instead of surface           */

/*
 * Plain-C mirror of the reference's `voxinfo` (src/util.h:50-69): byte-identical 64-byte layout
 * (bbox.min @0, bbox.max @12, gridsize @24, n_triangles @40, unit @48; alignment 8), so a
 * `const voxinfo*` can be passed where a `const voxb200_grid*` is expected.
 */
typedef struct voxb200_grid {
	float bbox_min[3];
	float bbox_max[3];
	unsigned int gridsize[3];
	size_t n_triangles;
	float unit[3];
} voxb200_grid;

/*
 * The part of the grid one call (one GPU) owns.  lo inclusive, hi exclusive, in voxel coordinates.
 * The table pointer passed alongside addresses ONLY this region: word 0 of it is the word holding
 * the region's first voxel.  The region must be contiguous in the chosen ordering:
 *   linear: full x and y range, any z range (a z-slab)          — G*G*(hi.z-lo.z)/8 bytes
 *   morton: a power-of-two aligned box produced by voxb200_partition (z halves, zy quadrants, octants…)
 * NULL means the whole grid.
 */
typedef struct voxb200_region {
	int lo[3];
	int hi[3];
} voxb200_region;

/* ---- device ------------------------------------------------------------------------------- */
int voxb200_device_count(int* count);
/* Selects `device` for the calling thread and checks it is a compute-capability-10.x part. */
int voxb200_init(int device);
const char* voxb200_last_error(void);

/* ---- grid parameters (host arithmetic, bit-identical to the reference's) -------------------- */
/* mesh_min/mesh_max: bbox over all mesh vertices.  Cube-ifies, pads by 1/10001, derives unit. */
int voxb200_make_grid(const float mesh_min[3], const float mesh_max[3], unsigned int gridsize,
                      size_t n_triangles, voxb200_grid* out);
/* Bytes of a whole table: max(the reference's ceil(G^3/32.0f)*4, the exact ceil(G^3/32)*4).  The two agree for every G that is a
 * multiple of 32 (all fast paths, morton order); for 629 odd sizes below 2048 the reference's binary32 division comes out a
 * word short and its own kernels write past the table (main.cpp:190) — size tables with THIS function.
 * voxb200_reference_table_bytes is the reference's number, for buffers that were sized by the reference's main(). */
size_t voxb200_table_bytes(unsigned int gridsize);
size_t voxb200_reference_table_bytes(unsigned int gridsize);
/* Region rank `part` of `n_parts` (n_parts a power of two <= 8 for morton; any n <= gridsize for
 * linear).  Writes the region and its size in bytes. */
int voxb200_partition(unsigned int gridsize, int morton, int part, int n_parts,
                      voxb200_region* out, size_t* region_bytes);
uint64_t voxb200_morton_encode(unsigned int x, unsigned int y, unsigned int z);

/* Frees every buffer, stream and event the library cached for the current device (work queue, staging, the persistent
 * buffers of voxb200_voxelize_host*).  The next call re-creates what it needs.  Synchronises the device. */
int voxb200_release(void);

/* ---- device memory ------------------------------------------------------------------------- */
int voxb200_malloc(void** dptr, size_t bytes);
int voxb200_free(void* dptr);
int voxb200_memcpy_d2h(void* host, const void* dptr, size_t bytes, void* stream);

/* ---- triangle upload ------------------------------------------------------------------------ */
/*
 * Host triangle soup (9 floats per triangle, model space) -> device.  soa4 = 0 keeps the
 * reference's AoS records (36 B/tri); soa4 = 1 transposes on the device into 3 float4 planes
 * (48 B/tri).  *d_tris is allocated here (voxb200_free it).  Asynchronous on `stream` apart from
 * the staging copy into pinned memory.
 */
int voxb200_upload_soup(const float* host_tris9, size_t n_triangles, int soa4, float** d_tris, void* stream);
/*
 * Indexed mesh (what trimesh2 holds: 12 B/vertex + 12 B/face) -> device triangles, expanded on
 * the GPU.  Also reduces the mesh bbox over all vertices on the device (trimesh2 need_bbox,
 * main.cpp:179) when mesh_min/mesh_max are non-NULL (synchronises the stream in that case).
 */
int voxb200_upload_indexed(const float* host_verts, size_t n_verts, const int32_t* host_faces, size_t n_faces,
                           int soa4, float** d_tris, float mesh_min[3], float mesh_max[3], void* stream);

/*
 * Upload-path option for a mesh that is voxelized more than once (per-frame use, benchmarks): a copy of the device soup
 * ordered by the z-layer of each triangle's lowest vertex (*d_sorted9 is allocated here, voxb200_free it).  The table
 * does not depend on the triangle order, the time does: with z-ordered triangles the atomics of voxb200_surface sweep
 * the table front to back and stay in L2 (10M triangles @2048^3: 0.65 -> 0.51 ms per voxelization; the sort itself
 * costs a few voxelizations).  Synchronises `stream`.  Part of meshToGPU_managed's replacement (main.cpp:61-80).
 */
int voxb200_sort_triangles(const voxb200_grid* grid, const float* d_tris9, float** d_sorted9, void* stream);

/*
 * Multi-GPU routing (new with the z-slab sharding; the reference is single-GPU): from a device soup, keep the
 * triangles that can touch `region` — surface: clamped grid bbox overlap (voxelize.cu:86-87); with
 * VOXB200_SOLID: centre-sample (y,z) range overlap (voxelize_solid.cu:112-113) — into a new compact device soup
 * (*d_routed, voxb200_free it).  Triangle order is not preserved (the table does not depend on it).
 * Synchronises `stream` to return the count.  Voxelizing the routed soup over `region` gives the same table
 * bytes as voxelizing the full soup over `region`.
 */
int voxb200_route_triangles(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* region,
                            float** d_routed, size_t* n_routed, void* stream);

/*
 * Routing to several regions at once (every rank routes ITS share of the mesh to all N ranks before an all-to-all):
 * d_out receives, back to back in region order, the triangles that can touch regions[0], regions[1], …;
 * counts[r] (host) = triangles in segment r; a triangle overlapping two regions appears in both segments.
 * out_capacity is in triangles.  n_regions <= 32.  Synchronises `stream`.
 */
int voxb200_route_triangles_multi(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* regions,
                                  int n_regions, float* d_out, size_t out_capacity, size_t* counts, void* stream);

/* ---- the hot path --------------------------------------------------------------------------- */
/*
 * d_tris and d_table are device-accessible pointers (cudaMalloc or cudaMallocManaged).  `stream`
 * is a cudaStream_t (NULL = legacy default stream).  Work is enqueued on `stream`; the call does
 * not synchronise.  Without VOXB200_ACCUMULATE the region's table bytes are zeroed first (that
 * write is part of the timed path).  Results are bit-identical to the reference's CPU voxelizer
 * (cpu_voxelizer.cpp) for the same voxinfo and triangles.
 */
int voxb200_surface(const voxb200_grid* grid, const float* d_tris, unsigned int* d_table,
                    unsigned int flags, const voxb200_region* region, void* stream);
int voxb200_solid(const voxb200_grid* grid, const float* d_tris, unsigned int* d_table,
                  unsigned int flags, const voxb200_region* region, void* stream);

/*
 * End to end with HOST buffers: upload the soup, voxelize (surface, or solid with VOXB200_SOLID),
 * copy the table back, synchronise.  host_table must hold voxb200_table_bytes(G) bytes (or the
 * region's bytes when region != NULL).  Fills timing_ms[0..3] (when non-NULL) with milliseconds: [0] H2D and [1]
 * voxelization (device events), [3] the whole call (host clock), [2] the rest: the table's way back.
 *
 * The way back (voxb200_download_table below): a dense copy, or — for a table of at least 32 MB whose non-zero words are
 * few (a surface table: 3 % on BASELINE config 4) and a 64-byte aligned host_table — the non-zero words only, as
 * {index, value} pairs in slices, expanded by host threads that stream the zero lines themselves while the next slice is on
 * the link (VOXB200_HOST_THREADS, default half the hardware threads, at most 16; VOXB200_READBACK=dense|sparse forces a
 * mode).  host_table is byte-identical to the device table either way.
 */
int voxb200_voxelize_host(const voxb200_grid* grid, const float* host_tris9, unsigned int* host_table,
                          unsigned int flags, const voxb200_region* region, float timing_ms[4]);

/*
 * Same, from the INDEXED mesh the caller holds (what the reference's main() has after TriMesh::read): uploads
 * 12 B/vertex + 12 B/face instead of the 36 B/triangle soup and expands on the GPU (main.cpp:61-80 replaced).
 * Face indices are range-checked on the device (out-of-range ones are clamped, counted, and fail the call with VOXB200_EINVAL).
 * grid->n_triangles = number of faces.
 */
int voxb200_voxelize_host_indexed(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                  unsigned int* host_table, unsigned int flags, const voxb200_region* region, float timing_ms[4]);

/*
 * The same call for a consumer that does not need the dense table: the table's NON-ZERO WORDS, ascending, as {index, bits} —
 * index = word of the (region's) table, bits = its 32 voxels, MSB first (voxel 32*index + k is bit 31-k).  Everything a writer
 * needs to walk the set voxels in table order, at 8 bytes per non-zero word over the link instead of the whole table (config 4:
 * 56 MB instead of 1 GiB) and without the host writing G^3/8 bytes.  *words points into a pinned buffer owned by the library,
 * valid until the next host entry point call on this device.  Tables of more than 2^32 words: voxelize in regions.
 */
typedef struct voxb200_word { uint32_t index; uint32_t bits; } voxb200_word;
int voxb200_voxelize_host_nonzero(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                  unsigned int flags, const voxb200_region* region, const voxb200_word** words, size_t* n_words, float timing_ms[4]);

/*
 * The read-back alone: d_table (table_words words on the current device, 16-byte aligned for the sparse mode) -> host_table,
 * synchronous, after everything enqueued on `stream`.  info (when non-NULL): [0] 1 = the sparse mode ran, [1] non-zero words.
 * Replaces the reference's reliance on managed memory for the writers' G^3 checkVoxel reads (main.cpp:229-253, util.h:25-38).
 */
int voxb200_download_table(const unsigned int* d_table, size_t table_words, unsigned int* host_table, void* stream, uint64_t info[2]);
/* What the last read-back of the current device's host entry points (voxb200_voxelize_host*, voxb200_download_table) did: same info[]. */
int voxb200_last_readback(uint64_t info[2]);
/* 0 = choose per table (default; VOXB200_READBACK=dense|sparse sets the initial value), 1 = always the dense copy, 2 = the sparse
 * mode whenever the pointers allow it.  Process-wide. */
int voxb200_set_readback_mode(int mode);
/* Host threads of the sparse read-back: 0 = default (VOXB200_HOST_THREADS, else half the hardware threads, at most 16).  Process-wide. */
int voxb200_set_host_threads(int n);
/* Diagnostics: exercises the read-back's host-thread machinery (no GPU needed).  0 = ok. */
int voxb200_selftest_host_pool(void);

/* ---- multi-GPU: one process, one host thread per device ------------------------------------------------------ */
/*
 * The reference's caller is a single-threaded main() that holds the indexed mesh and wants the table in host memory
 * (main.cpp:203-222).  This serves that caller with n_devices GPUs of one box (devices[] = CUDA ordinals, NULL = 0..n-1):
 * every device copies 1/N of the mesh bytes over its own PCIe link, the shares are all-gathered device to device
 * (cudaMemcpyPeerAsync: NVLink where peer access exists), device d voxelizes region d of voxb200_partition(G, morton, d, N) —
 * disjoint slices of the table, no reduction — and copies its slab straight into its byte range of host_table
 * (voxb200_table_bytes(G) bytes, pinned for full speed: voxb200_host_alloc).  Synchronous.  flags: VOXB200_SOLID, VOXB200_MORTON.
 * The table is bit-identical to the single-GPU table.  timing_ms (when non-NULL), device milliseconds, maximum over the
 * devices: [0] H2D of the shares, [1] peer all-gather, [2] preparation (tile records / expansion), [3] voxelization,
 * [4] the slab's way back (host clock: dense copy or sparse read-back, see voxb200_voxelize_host), [5] first H2D byte to the
 * slab complete in host memory; [6] host wall-clock of the call, [7] n_devices.
 * Linear order needs G*G divisible by 32 for N > 1; morton order a power-of-two N <= 8.
 */
int voxb200_voxelize_host_multi(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                unsigned int* host_table, unsigned int flags, const int* devices, int n_devices, float timing_ms[8]);
/* Device-resident slabs (slab k on device slab_devices[k], slab_bytes[k] bytes, in region order) -> one table on table_device:
 * peer copies enqueued on `stream` (a stream of the current device).  The NCCL-free slab gather of SURVEY §8e. */
int voxb200_gather_slabs(unsigned int* const* d_slabs, const int* slab_devices, const size_t* slab_bytes, int n_slabs,
                         unsigned int* d_table, int table_device, void* stream);
/* Pinned (page-locked, portable) host memory for the host entry points. */
int voxb200_host_alloc(void** host_ptr, size_t bytes);
int voxb200_host_free(void* host_ptr);

/* ---- prepared meshes: the resident / per-frame interface ------------------------------------------------ */
/*
 * The reference voxelizes a mesh once per process (main.cpp:203-222) although it pitches per-frame use (README.md:74).
 * A voxb200_mesh is a mesh prepared for ONE grid (and region): it owns a re-ordered copy of the triangles, the plan of
 * the tile-owner surface schedule and a private workspace, so
 *   - voxb200_mesh_voxelize only enqueues kernels on `stream` (no allocation, no synchronisation, capturable into a
 *     CUDA graph) and is RE-ENTRANT: different meshes may voxelize concurrently on different streams of a device;
 *   - a surface voxelization in linear order on a grid whose size is a multiple of 256 (<= 4096) writes every table
 *     byte exactly once: tiles of 256x16x16 voxels are assembled in shared memory by the one thread block that owns
 *     them, the empty tiles are cleared on the side — no zero-fill pass, no global atomics except for the triangles
 *     larger than 4x4x4 voxels, which take the row-solver path afterwards.  Every other configuration (solid, morton
 *     order, other grid sizes) runs the one-shot kernels on the handle's own data.
 * The table is bit-identical to voxb200_surface / voxb200_solid on the same triangles.
 * flags of _create: VOXB200_SOLID, VOXB200_MORTON.  flags of _voxelize: VOXB200_ACCUMULATE.
 * _create / _update synchronise `stream` (they size buffers from device-side counts); _update re-prepares the mesh for new
 * vertex positions (same triangle count) reusing the handle's buffers.  The triangle arrays are not referenced after
 * the call returns.  voxb200_mesh_info: [0] 1 = tile schedule, [1] tiles, [2] non-empty tiles, [3] binned triangle
 * instances, [4] triangles on the row-solver side path, [5] 32-triangle batches, [6] zero-fill chunks per batch,
 * [7] zero-only blocks.
 */
typedef struct voxb200_mesh voxb200_mesh;
int voxb200_mesh_create(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* region,
                        voxb200_mesh** out, void* stream);
int voxb200_mesh_create_indexed(const voxb200_grid* grid, const float* d_verts, size_t n_verts, const int32_t* d_faces, unsigned int flags,
                                const voxb200_region* region, voxb200_mesh** out, void* stream);
int voxb200_mesh_update(voxb200_mesh* mesh, const float* d_tris9, void* stream);
int voxb200_mesh_update_indexed(voxb200_mesh* mesh, const float* d_verts, size_t n_verts, const int32_t* d_faces, void* stream);
int voxb200_mesh_voxelize(voxb200_mesh* mesh, unsigned int* d_table, unsigned int flags, void* stream);
int voxb200_mesh_info(const voxb200_mesh* mesh, uint64_t out[8]);
/* voxb200_last_counters for the mesh's last voxelization (its private workspace); synchronises the device. */
int voxb200_mesh_counters(const voxb200_mesh* mesh, uint64_t out[4]);
int voxb200_mesh_destroy(voxb200_mesh* mesh);

/*
 * Table consumer (replaces the G^3 host checkVoxel() loops of the writers, src/util_io.cpp:92-285): compacts the set
 * bits of `table_words` table words into a device array of voxel indices, ascending — linear idx = x + G*y + G*G*z
 * (or the morton code for a morton table); `first_voxel` is added to every index (the index of a region's first
 * voxel, 0 for a whole table).  *d_indices is allocated here (voxb200_free it).  Synchronises `stream`.
 */
int voxb200_extract_voxels(const unsigned int* d_table, size_t table_words, uint64_t first_voxel, uint64_t** d_indices, size_t* count, void* stream);
/*
 * Table consumer for the binvox writer (util_io.cpp:202-246): the run-length encoded payload — (value, count <= 255) byte pairs
 * over the voxels visited x-major, then z, then y — of a whole LINEAR table, built on the device, byte-identical to what the
 * reference's G^3 checkVoxel loop writes after its ASCII header.  gridsize: a multiple of 256, at most 4096 (smaller grids: walk
 * the voxel list).  *d_bytes is allocated here (voxb200_free it), *n_bytes its size.  Synchronises `stream`.
 */
int voxb200_binvox_rle(const unsigned int* d_table, unsigned int gridsize, unsigned char** d_bytes, size_t* n_bytes, void* stream);

/* ---- introspection ---------------------------------------------------------------------------- */
/* Kernels launched by this library since the last reset (the bench's "gpu_launches"). */
uint64_t voxb200_launch_count(int reset);
/* Counters of the last surface/solid call, valid after the stream has been synchronised:
 * [0] triangles routed to the cooperative (large-triangle) path, [1] work units of that path ((y,z) rows for the surface
 * path, blocks of 256 centre samples for the solid path; UINT64_MAX when more than 2^32 units were queued in one call — the
 * table contents are undefined then; voxb200_voxelize_host* and the C++ drop-in symbols report it as an error),
 * [2] solid: samples clamped because xmax fell outside [0, G-1] (reference UB territory),
 * [3] 1 when the last voxb200_solid call used per-row mark lists + a single fill pass (no zero-fill, no scan) — host-side state. */
int voxb200_last_counters(uint64_t out[4]);
/*
 * Per-phase device timing.  With profiling on, every voxb200_surface / voxb200_solid call records CUDA
 * events on its own stream between its kernels (a ring of the last 256 calls).  After the stream has
 * been synchronised, voxb200_phase_ms(i, out) gives, for the i-th call since profiling was enabled,
 * milliseconds of: [0] table zero-fill, [1] per-triangle kernel, [2] cooperative (large-triangle)
 * kernel, [3] solid column scan (0 for surface).
 */
int voxb200_set_profiling(int on);
int voxb200_phase_ms(unsigned int call_index, float out[4]);
const char* voxb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VOXB200_H */
