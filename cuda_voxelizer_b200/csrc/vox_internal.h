// vox_internal.h — launcher interface between the C ABI (vox_abi.cu) and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <utility>
#include "vox_exact.cuh"

namespace voxb {

constexpr int kProfRing = 256;     // calls remembered
constexpr int kProfEvents = 5;     // start | zero done | per-triangle kernel done | cooperative kernel done | scan done

// Device scratch owned by the library, one per device, grown on demand (never shrunk).
struct Workspace {
	int device = -1;
	int sm_count = 0;
	unsigned long long* counters = nullptr;   // kNumCounters x u64, zeroed at the start of every call
	uint2* queue = nullptr;                   // cooperative-path queue: {triangle, first work item}
	size_t queue_cap = 0;                     // entries
	uint4* setups = nullptr;                  // stored SurfSetup of the first setup_cap queued triangles (kSetupVec x 16 B each)
	size_t setup_cap = 0;
	unsigned int* dir = nullptr;              // work-item directory (one slot per 64 items)
	size_t dir_cap = 0;
	QueueView view() const {
		QueueView q;
		q.entries = queue; q.cursor = counters + 0; q.setups = setups; q.setup_cap = (unsigned int)setup_cap;
		q.dir = dir; q.dir_cap = (unsigned int)dir_cap;
		return q;
	}
	unsigned int* route_masks = nullptr;      // multi-region routing: per-triangle region mask
	size_t route_cap = 0;
	unsigned long long* route_counts = nullptr;   // 32 counters + 32 cursors
	unsigned int* scratch = nullptr;          // solid: mark table for ACCUMULATE / morton modes; overflow table of the row-list mode
	size_t scratch_words = 0;
	bool scratch_zero = false;                // scratch is all-zero (the row-list mode keeps it so between calls)
	unsigned int* row_count = nullptr;        // solid row lists: marks per (y,z) row (all-zero between calls) ...
	unsigned short* row_marks = nullptr;      // ... and kRowMarks 16-bit xmax slots per row
	size_t row_cap = 0;                       // rows
	bool rows_dirty = false;                  // a call failed between the mark and fill phases: counters / slots / spill table are re-initialised next time
	bool last_row_lists = false;              // the last solid call took the row-list schedule (voxb200_last_counters()[3])
	// optional per-phase timing (voxb200_set_profiling): a ring of event sets, one set per call
	bool prof_on = false;
	unsigned int prof_calls = 0;
	cudaEvent_t (*prof_ev)[kProfEvents] = nullptr;    // [kProfRing][kProfEvents]
};

// Records phase boundary `which` of the current call on `st` when profiling is on.
inline void prof_mark(Workspace& ws, int which, cudaStream_t st) {
	if (ws.prof_on && ws.prof_ev) cudaEventRecord(ws.prof_ev[ws.prof_calls % kProfRing][which], st);
}

enum Counter {
	kCtrQueue = 0,        // packed: (queue entries << 32) | work items
	kCtrClaim = 1,        // next work item to hand out (dynamic scheduling)
	kCtrSolidClamp = 2,   // solid samples whose xmax fell outside [0, G-1]
	kCtrQueueOverflow = 3,   // the work-unit half of kCtrQueue wrapped (more than 2^32 queued rows / sample blocks in one call)
	kNumCounters = 8
};

struct LaunchOpts {
	bool morton;
	bool accumulate;
	bool soa4;
};

#ifdef __CUDACC__
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining; it calls
// grid_dependency_wait() before it touches anything the predecessor wrote, so only its launch latency and block ramp overlap the
// predecessor's tail (the three short kernels of a solid voxelization: a few microseconds each).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// "my dependents may be scheduled": issued at the top of a block — dependents are launched once EVERY block of this grid has issued it
// (or exited), i.e. during this grid's last wave, and then sit in grid_dependency_wait() until this grid has completed and flushed.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), unsigned int blocks, unsigned int threads, cudaStream_t st, Args&&... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#endif

extern std::atomic<unsigned long long> g_launch_count;   // kernels launched by this library (voxb200_launch_count); device threads of the multi-device call bump it concurrently

cudaError_t ensure_queue(Workspace& ws, size_t entries);
cudaError_t ensure_scratch(Workspace& ws, size_t words);
// Row-list buffers for n_rows rows plus a ZEROED scratch (overflow) table of `words` words.
cudaError_t ensure_row_lists(Workspace& ws, size_t n_rows, size_t words, cudaStream_t st);

// Per-row mark lists of the solid path (solid.cu)
struct RowLists {
	unsigned int* count;
	unsigned short* marks;
};

// Zeroes `words` 32-bit words at p (own kernel: 16-byte stores, grid sized to the SM count).
cudaError_t launch_zero(Workspace& ws, unsigned int* p, size_t words, cudaStream_t st, bool reset_counters = false);

// The surface path (voxelize.cu:58-238 replaced): [zero] + per-triangle kernel + cooperative kernel.
cudaError_t launch_surface(Workspace& ws, const GridParams& g, const float* d_tris, unsigned int* d_table,
                           size_t region_words, const LaunchOpts& o, cudaStream_t st);
// The solid path (voxelize_solid.cu:73-193 replaced): [zero] + mark kernels + column suffix-XOR scan.
cudaError_t launch_solid(Workspace& ws, const GridParams& g, const float* d_tris, unsigned int* d_table,
                         size_t region_words, const LaunchOpts& o, cudaStream_t st);

// Upload helpers (main.cpp:61-80 replaced)
cudaError_t launch_soup_to_soa4(const float* d_soup, float* d_soa4, size_t n_tris, cudaStream_t st);
cudaError_t launch_expand_indexed(const float* d_verts, const int* d_faces, size_t n_faces, size_t n_verts,
                                  bool soa4, float* d_out, cudaStream_t st, unsigned long long* d_bad = nullptr);      // n_verts != 0: indices outside [0, n_verts) are clamped and counted into *d_bad
// Upload-path layer sort: d_out = d_soup ordered by the z-layer of each triangle's lowest vertex (d_keys: n_tris words, d_hist: G words of scratch)
cudaError_t launch_layer_sort(const GridParams& g, const float* d_soup, float* d_out, unsigned int* d_keys, unsigned int* d_hist, cudaStream_t st);
cudaError_t launch_route(const GridParams& g, bool solid, const float* d_soup, float* d_out, unsigned long long* d_cursor, cudaStream_t st);
cudaError_t launch_route_count(const GridParams& g, bool solid, const int (*lo)[3], const int (*hi)[3], int n_regions, const float* d_soup,
                               unsigned int* d_masks, unsigned long long* d_counts, cudaStream_t st);
cudaError_t launch_route_scatter(unsigned long long n_tris, int n_regions, const float* d_soup, const unsigned int* d_masks, float* d_out,
                                 unsigned long long* d_cursors, cudaStream_t st);
// ---- tile-owner surface path of a prepared mesh (tiles.cu) ------------------------------------------------
// A tile is Tx x kTileY x kTileZ voxels with Tx = min(G, 1024): its row segments are whole 128-byte lines (or whole rows), so the
// owning block's clears and atomics never share an L2 line with another block.
#ifndef VOXB_TILE_Y
#define VOXB_TILE_Y 16
#endif
#ifndef VOXB_TILE_Z
#define VOXB_TILE_Z 16
#endif
constexpr int kTileY = VOXB_TILE_Y, kTileZ = VOXB_TILE_Z;
constexpr int kTileXMax = 1024;
constexpr int kZeroBlockChunks = 256;                      // 512-byte chunks per zero-only block of the tile kernel (128 KB)
struct TileGeom {
	int ntx, nty, ntz;              // tiles per axis inside the region
	int tx_shift;                   // log2 of the tile's x-extent in voxels (8..10)
	int chunk_shift;                // log2 of the 512-byte zero-fill chunks per tile
	int G;
	int tz0;                        // first tile layer of the region (region z0 / kTileZ)
	unsigned int n_tiles;
	unsigned int n_verts;           // indexed input: vertices (0: not known, indices are trusted)
};
enum PlanTotal {                    // device-side totals of the planning pass (u64 each)
	kPlanInstances = 0,             // (triangle, tile) records
	kPlanEmpty = 1, kPlanWork = 2,  // empty / non-empty tiles
	kPlanBatches = 3,               // 32-triangle batches over all non-empty tiles
	kPlanBigDirect = 4,             // triangles whose bbox exceeds 4x4x4 voxels
	kPlanHeavyInstances = 5,        // instances that fell into over-full tiles (they take the side path)
	kPlanSideFill = 6,              // triangles written to the side soup
	kPlanWide = 7,                  // small triangles that need the 64-candidate evaluation
	kPlanBadFaces = 8,              // faces with a vertex index outside [0, n_verts) (indexed input; they are clamped and reported)
	kPlanTotals = 9
};
struct TilePlan {
	TileGeom geom;
	const float* soup;              // 64-byte records (tiles.cu), one per (triangle, tile); the run of tile t starts at record off[t]
	const unsigned int* cnt;        // triangles per tile
	const unsigned int* off;
	const unsigned int* order;      // non-empty tiles, heaviest first (n_work)
	const uint4* work;              // per non-empty tile, heaviest first (n_work): {table word of its first voxel, records, first record, batches before it}
	const unsigned int* empty;      // empty tiles in table order (n_empty): the table word, relative to the region, of each one's first voxel
	unsigned int n_work, n_empty;
	unsigned int zero_chunks;       // 16 * n_empty chunks of 512 bytes
	unsigned int zero_quota;        // chunks every 32-triangle batch clears
	unsigned int zero_rest_first, n_zero_blocks;      // chunks from here on are cleared by zero-only blocks
	bool wide;                      // some binned triangle needs the <=4x4x4 evaluation
};
cudaError_t launch_tile_count(const GridParams& g, const TileGeom& tg, const float* d_soup, const float* d_verts, const int* d_faces,
                              unsigned int* d_keys, unsigned int* d_cnt, unsigned long long* d_totals, cudaStream_t st);
cudaError_t launch_tile_plan(const TileGeom& tg, unsigned int cap, const unsigned int* d_cnt, unsigned int* d_off, unsigned int* d_order,
                             void* d_work, unsigned int* d_empty, unsigned long long* d_totals, unsigned int* d_scratch, cudaStream_t st);      // d_scratch: n_tiles words
cudaError_t launch_tile_scatter(const GridParams& g, const TileGeom& tg, unsigned int cap, const float* d_soup, const float* d_verts,
                                const int* d_faces, const unsigned int* d_keys, const unsigned int* d_cnt, const unsigned int* d_off,
                                unsigned int* d_fill, void* d_records, float* d_side, unsigned long long* d_totals, cudaStream_t st);
bool mesh_tileable(const GridParams& g, unsigned int flags);      // mesh.cu: the tile schedule covers this grid / region / mode
cudaError_t launch_surface_tiles(const GridParams& g, const TilePlan& p, unsigned int* d_table, bool accumulate, cudaStream_t st);

// ---- shared by the ABI translation units (defined in vox_abi.cu) ---------------------------------------------
int abi_fail(int code, const char* fmt, ...);
int abi_fail_cuda(cudaError_t e, const char* what);
int abi_current_ws(Workspace** out);                 // the calling thread's current device's library workspace
int abi_init_workspace(Workspace& ws, int dev);      // a private workspace (prepared meshes own one each)
void abi_free_workspace(Workspace& ws);
void multi_release_device(int dev);                 // multi.cu
}  // namespace voxb
struct voxb200_grid;
struct voxb200_region;
namespace voxb {
int abi_resolve_region(const ::voxb200_grid* grid, const ::voxb200_region* region, bool morton, GridParams* g, size_t* region_words);

// Table consumer: set voxels -> ascending voxel indices
size_t extract_blocks(size_t n_words);
cudaError_t launch_extract_count(const unsigned int* d_table, size_t n_words, unsigned int* d_counts, unsigned long long* d_offsets, cudaStream_t st);
cudaError_t launch_extract_write(const unsigned int* d_table, size_t n_words, const unsigned long long* d_offsets, unsigned long long first_voxel,
                                 unsigned long long* d_out, cudaStream_t st);
// Non-zero words -> ascending {word index, value} pairs (extract.cu); blocks as extract_blocks(), 2048 words each
constexpr size_t kNzBlockWords = 2048;
cudaError_t launch_nz_count(const unsigned int* d_table, size_t n_words, unsigned int* d_counts, unsigned long long* d_offsets, cudaStream_t st);
cudaError_t launch_nz_write(const unsigned int* d_table, size_t n_words, const unsigned long long* d_offsets, void* d_pairs, cudaStream_t st);

// Device table -> host table (readback.cu): a dense copy, or — when few words are non-zero — the non-zero words only, expanded
// by host threads that stream the zeros themselves.  Per-device persistent buffers.
struct Readback {
	unsigned int* d_counts = nullptr; unsigned long long* d_offsets = nullptr; size_t blocks_cap = 0;
	unsigned long long* h_offsets = nullptr; size_t h_blocks_cap = 0;       // pinned
	void* d_pairs = nullptr; size_t pairs_cap = 0;                          // pairs
	void* h_pairs = nullptr; size_t h_pairs_cap = 0;                        // pinned
	cudaEvent_t* ev = nullptr; int n_ev = 0;
	cudaEvent_t go_ev = nullptr;            // recorded behind the upload: the zero-fill workers that wait for it poll it
	void* host = nullptr;                   // host threads + the zero-fill that runs ahead (readback.cu)
	// what the last call did
	int last_mode = 0;                      // 0 dense copy, 1 sparse
	unsigned long long last_nonzero = 0;    // non-zero words (sparse mode)
};
int readback_table(Readback& rb, const unsigned int* d_table, size_t words, unsigned int* host_table, cudaStream_t st, int host_threads);
// Optional, in a call whose table is expected to be sparse: host threads zero-fill host_table ahead of the read-back, starting when
// `after` has passed the point it is at now (enqueue it behind the upload's copies).
void readback_prezero(Readback& rb, unsigned int* host_table, size_t words, int host_threads, cudaStream_t after, bool allow_early);
void readback_cancel(Readback& rb);          // every exit of a call that started one (idempotent)
struct ReadbackGuard { Readback& rb; ~ReadbackGuard() { readback_cancel(rb); } };
int readback_pairs(Readback& rb, const unsigned int* d_table, size_t words, cudaStream_t st, const void** host_pairs, size_t* n_pairs);
void readback_free(Readback& rb);
int readback_default_threads();
cudaError_t launch_bbox_reduce(const float* d_verts, size_t n_verts, float* d_minmax6, cudaStream_t st);

}  // namespace voxb
