// vox_abi.cu — the C ABI declared in include/voxb200.h: device selection, grid parameters, triangle
// upload, region bookkeeping and the launch of the surface / solid paths.  No CPU fallback lives here:
// every compute entry point fails with VOXB200_ENODEVICE when there is no sm_100 device.
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/voxb200.h"
#include "vox_internal.h"

using namespace voxb;

namespace {

constexpr int kMaxDevices = 64;
Workspace g_ws[kMaxDevices];
thread_local char g_err[512] = "";

// e2e scratch of voxb200_voxelize_host, per device, grown on demand
struct HostPath {
	float* d_tris = nullptr; size_t tris_bytes = 0;
	unsigned int* d_table = nullptr; size_t table_bytes = 0;
	float* d_verts = nullptr; size_t verts_bytes = 0;
	int* d_faces = nullptr; size_t faces_bytes = 0;
	void* pinned[2] = {nullptr, nullptr}; size_t pinned_bytes = 0;
	voxb200_mesh* mesh = nullptr;           // prepared mesh of voxb200_voxelize_host_indexed (tile schedule), re-prepared per call
	voxb200_grid mesh_grid{}; voxb200_region mesh_region{}; bool mesh_has_region = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t buf_free[2] = {nullptr, nullptr};
	bool buf_busy[2] = {false, false};     // a copy out of pinned[b] may still be in flight
	int next_buf = 0;
	Readback rb;                            // device table -> host table (readback.cu)
	bool last_tiles = false;                // the last indexed host call took the tile schedule
	unsigned long long* d_bad = nullptr;    // faces with an out-of-range vertex index seen by the expansion of the last indexed host call
};
HostPath g_hp[kMaxDevices];

int fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
int fail_cuda(cudaError_t e, const char* what) {
	int code = (e == cudaErrorMemoryAllocation) ? VOXB200_ENOMEM
	         : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice) ? VOXB200_ENODEVICE
	         : VOXB200_ECUDA;
	return fail(code, "%s: CUDA error %d (%s) \"%s\"", what, (int)e, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(e_, #call); } while (0)

// Workspace of the current device (initialising it on first use).
int current_ws(Workspace** out) {
	int dev = -1;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
	if (dev < 0 || dev >= kMaxDevices) return fail(VOXB200_ENODEVICE, "device ordinal %d out of range", dev);
	Workspace& ws = g_ws[dev];
	if (ws.device != dev) {
		cudaDeviceProp prop;
		CU(cudaGetDeviceProperties(&prop, dev));
		if (prop.major != 10)
			return fail(VOXB200_ENODEVICE, "device %d (%s) is compute capability %d.%d; this library is built for sm_100a only",
			            dev, prop.name, prop.major, prop.minor);
		CU(cudaMalloc(&ws.counters, kNumCounters * sizeof(unsigned long long)));
		CU(cudaMemset(ws.counters, 0, kNumCounters * sizeof(unsigned long long)));
		ws.sm_count = prop.multiProcessorCount;
		ws.device = dev;
	}
	*out = &ws;
	return VOXB200_OK;
}

// floor(log2(v)) for v > 0
int ilog2(unsigned long long v) { int r = 0; while (v >>= 1) r++; return r; }
unsigned int compact3(unsigned long long m) {   // inverse of spread3
	unsigned int r = 0;
	for (int i = 0; i < 21; i++) r |= (unsigned int)((m >> (3 * i)) & 1ull) << i;
	return r;
}

// Validates the region and derives {word_base, region_words}.
static int resolve_region_impl(const voxb200_grid* grid, const voxb200_region* region, bool morton, GridParams* g, size_t* region_words) {
	const unsigned int G = grid->gridsize[0];
	if (G == 0 || grid->gridsize[1] != G || grid->gridsize[2] != G)
		return fail(VOXB200_EINVAL, "grid must be cubic and non-empty (got %u %u %u); the reference only builds cubic grids (main.cpp:186)",
		            grid->gridsize[0], grid->gridsize[1], grid->gridsize[2]);
	if (G > 65536u) return fail(VOXB200_EINVAL, "gridsize %u too large", G);
	const unsigned long long G3 = (unsigned long long)G * G * G;
	g->bx = grid->bbox_min[0]; g->by = grid->bbox_min[1]; g->bz = grid->bbox_min[2];
	g->ux = grid->unit[0]; g->uy = grid->unit[1]; g->uz = grid->unit[2];
	g->rux = 1.0f / g->ux; g->ruy = 1.0f / g->uy; g->ruz = 1.0f / g->uz;
	g->G = (int)G;
	g->n_tris = grid->n_triangles;
	if (grid->n_triangles > 0xfffffff0ull) return fail(VOXB200_EINVAL, "more than 2^32 triangles in one call");
	if (morton && (G & (G - 1)) != 0) return fail(VOXB200_EINVAL, "morton order needs a power-of-two grid size (got %u)", G);
	if (!region) {
		g->rx0 = g->ry0 = g->rz0 = 0;
		g->rx1 = g->ry1 = g->rz1 = (int)G;
		g->word_base = 0;
		*region_words = voxb200_table_bytes(G) / 4;
		return VOXB200_OK;
	}
	for (int k = 0; k < 3; k++)
		if (region->lo[k] < 0 || region->hi[k] > (int)G || region->lo[k] >= region->hi[k])
			return fail(VOXB200_EINVAL, "empty or out-of-grid region on axis %d: [%d, %d)", k, region->lo[k], region->hi[k]);
	g->rx0 = region->lo[0]; g->rx1 = region->hi[0];
	g->ry0 = region->lo[1]; g->ry1 = region->hi[1];
	g->rz0 = region->lo[2]; g->rz1 = region->hi[2];
	const bool whole = g->rx0 == 0 && g->ry0 == 0 && g->rz0 == 0 && g->rx1 == (int)G && g->ry1 == (int)G && g->rz1 == (int)G;
	if (whole) { g->word_base = 0; *region_words = voxb200_table_bytes(G) / 4; return VOXB200_OK; }
	if (!morton) {
		if (g->rx0 != 0 || g->rx1 != (int)G || g->ry0 != 0 || g->ry1 != (int)G)
			return fail(VOXB200_EINVAL, "a linear-order region must span the full x and y range (z-slab)");
		const unsigned long long first = (unsigned long long)G * G * (unsigned long long)g->rz0;
		const unsigned long long last = (unsigned long long)G * G * (unsigned long long)g->rz1;
		if ((first & 31ull) || ((last & 31ull) && last != G3))
			return fail(VOXB200_EINVAL, "z-slab [%d,%d) of a %u^3 grid does not start/end on a table word", g->rz0, g->rz1, G);
		g->word_base = first >> 5;
		*region_words = (size_t)(((last + 31ull) >> 5) - g->word_base);
		return VOXB200_OK;
	}
	const unsigned int sx = g->rx1 - g->rx0, sy = g->ry1 - g->ry0, sz = g->rz1 - g->rz0;
	if ((sx & (sx - 1)) || (sy & (sy - 1)) || (sz & (sz - 1)) || (g->rx0 & (sx - 1)) || (g->ry0 & (sy - 1)) || (g->rz0 & (sz - 1)))
		return fail(VOXB200_EINVAL, "a morton-order region must be a power-of-two box aligned to its own size");
	const int a = ilog2(sx), b = ilog2(sy), c = ilog2(sz);
	if (!(a >= b && b >= c && c >= a - 1) || a + b + c < 5)
		return fail(VOXB200_EINVAL, "box %ux%ux%u is not a contiguous run of the morton curve", sx, sy, sz);
	const unsigned long long first = morton3((unsigned)g->rx0, (unsigned)g->ry0, (unsigned)g->rz0);
	g->word_base = first >> 5;
	*region_words = (size_t)(1ull << (a + b + c - 5));
	return VOXB200_OK;
}
// + w32: region-relative word offsets fit 32 bits (the 32-bit addressing of the scatter and tile code)
int resolve_region(const voxb200_grid* grid, const voxb200_region* region, bool morton, GridParams* g, size_t* region_words) {
	g->w32 = 0;
	const int rc = resolve_region_impl(grid, region, morton, g, region_words);
	if (rc == VOXB200_OK) g->w32 = *region_words <= 0x100000000ull ? 1 : 0;
	return rc;
}

int run_path(bool solid, const voxb200_grid* grid, const float* d_tris, unsigned int* d_table, unsigned int flags,
             const voxb200_region* region, cudaStream_t st) {
	if (!grid || !d_table || (!d_tris && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL grid / triangle / table pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	GridParams g;
	size_t region_words = 0;
	const bool morton = (flags & VOXB200_MORTON) != 0;
	rc = resolve_region(grid, region, morton, &g, &region_words);
	if (rc) return rc;
	LaunchOpts o;
	o.morton = morton;
	o.accumulate = (flags & VOXB200_ACCUMULATE) != 0;
	o.soa4 = (flags & VOXB200_TRIS_SOA4) != 0;
	if (o.soa4 && (reinterpret_cast<uintptr_t>(d_tris) & 15u))
		return fail(VOXB200_EINVAL, "SoA float4 triangle planes must be 16-byte aligned");
	cudaError_t e = solid ? launch_solid(*ws, g, d_tris, d_table, region_words, o, st)
	                      : launch_surface(*ws, g, d_tris, d_table, region_words, o, st);
	if (e != cudaSuccess) return fail_cuda(e, solid ? "voxb200_solid launch" : "voxb200_surface launch");
	return VOXB200_OK;
}

bool is_pinned_or_device(const void* p) {
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged || attr.type == cudaMemoryTypeDevice;
}

// Grows a persistent device buffer of the end-to-end path.
template <typename T>
int grow(T** p, size_t* have, size_t want) {
	if (want <= *have) return VOXB200_OK;
	if (*p) cudaFree(*p);
	*p = nullptr; *have = 0;
	CU(cudaMalloc(p, want));
	*have = want;
	return VOXB200_OK;
}

constexpr size_t kStageBytes = 32u << 20;   // pinned staging chunk for pageable host buffers

int ensure_host_path(HostPath& hp) {
	if (!hp.stream) {
		CU(cudaStreamCreateWithFlags(&hp.stream, cudaStreamNonBlocking));
		for (auto& e : hp.ev) CU(cudaEventCreate(&e));
		for (auto& e : hp.buf_free) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	return VOXB200_OK;
}
int ensure_pinned(HostPath& hp) {
	if (!hp.pinned[0]) {
		CU(cudaMallocHost(&hp.pinned[0], kStageBytes));
		CU(cudaMallocHost(&hp.pinned[1], kStageBytes));
		hp.pinned_bytes = kStageBytes;
	}
	return VOXB200_OK;
}

// host -> device on `st`: direct async copy when the source is pinned, else a double-buffered
// pageable -> pinned -> device pipeline (what "pinned staging + cudaMemcpyAsync" means in SURVEY §7.4).
int h2d(HostPath& hp, void* dst, const void* src, size_t bytes, cudaStream_t st) {
	if (bytes == 0) return VOXB200_OK;
	if (is_pinned_or_device(src)) { CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st)); return VOXB200_OK; }
	int rc = ensure_pinned(hp);
	if (rc) return rc;
	size_t off = 0;
	while (off < bytes) {
		const int b = hp.next_buf;
		const size_t n = bytes - off < hp.pinned_bytes ? bytes - off : hp.pinned_bytes;
		if (hp.buf_busy[b]) CU(cudaEventSynchronize(hp.buf_free[b]));   // also across calls: the staging buffers are shared
		memcpy(hp.pinned[b], (const char*)src + off, n);
		CU(cudaMemcpyAsync((char*)dst + off, hp.pinned[b], n, cudaMemcpyHostToDevice, st));
		CU(cudaEventRecord(hp.buf_free[b], st));
		hp.buf_busy[b] = true;
		off += n;
		hp.next_buf ^= 1;
	}
	return VOXB200_OK;
}

}  // namespace

namespace voxb {
int abi_fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
int abi_fail_cuda(cudaError_t e, const char* what) { return fail_cuda(e, what); }
int abi_current_ws(Workspace** out) { return current_ws(out); }
int abi_resolve_region(const voxb200_grid* grid, const voxb200_region* region, bool morton, GridParams* g, size_t* region_words) {
	return resolve_region(grid, region, morton, g, region_words);
}
int abi_init_workspace(Workspace& ws, int dev) {
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, dev));
	CU(cudaMalloc(&ws.counters, kNumCounters * sizeof(unsigned long long)));
	CU(cudaMemset(ws.counters, 0, kNumCounters * sizeof(unsigned long long)));
	ws.sm_count = prop.multiProcessorCount;
	ws.device = dev;
	return VOXB200_OK;
}
void abi_free_workspace(Workspace& ws) {
	void* dev_ptrs[] = {ws.counters, ws.queue, ws.setups, ws.dir, ws.route_masks, ws.route_counts, ws.scratch, ws.row_count, ws.row_marks};
	for (void* p : dev_ptrs) if (p) cudaFree(p);
	if (ws.prof_ev) {
		for (int i = 0; i < kProfRing; i++) for (int k = 0; k < kProfEvents; k++) cudaEventDestroy(ws.prof_ev[i][k]);
		delete[] ws.prof_ev;
	}
	ws = Workspace();
}
cudaError_t ensure_queue(Workspace& ws, size_t entries) {
	if (entries <= ws.queue_cap) return cudaSuccess;
	if (ws.queue) cudaFree(ws.queue);
	ws.queue = nullptr; ws.queue_cap = 0;
	size_t cap = entries + entries / 8 + 1024;
	cudaError_t e = cudaMalloc(&ws.queue, cap * sizeof(uint2));
	if (e != cudaSuccess) return e;
	ws.queue_cap = cap;
	// setups of queued triangles: bounded (160 B each); queue slots beyond the cap recompute their setup instead
	const size_t want = cap < (size_t(1) << 21) ? cap : (size_t(1) << 21);
	if (want > ws.setup_cap) {
		if (ws.setups) cudaFree(ws.setups);
		ws.setups = nullptr; ws.setup_cap = 0;
		if (cudaMalloc(&ws.setups, want * kSetupVec * sizeof(uint4)) == cudaSuccess) ws.setup_cap = want;
		else cudaGetLastError();          // optional: without it every slot recomputes
	}
	if (!ws.dir) {
		const size_t dir_entries = size_t(1) << 20;          // covers 64 M work items; items beyond it fall back to the full search
		if (cudaMalloc(&ws.dir, dir_entries * sizeof(unsigned int)) == cudaSuccess) ws.dir_cap = dir_entries;
		else cudaGetLastError();
	}
	return cudaSuccess;
}
cudaError_t ensure_row_lists(Workspace& ws, size_t n_rows, size_t words, cudaStream_t st) {
	cudaError_t e = ensure_scratch(ws, words);
	if (e != cudaSuccess) return e;
	if (ws.rows_dirty) ws.scratch_zero = false;
	if (!ws.scratch_zero) {
		e = launch_zero(ws, ws.scratch, ws.scratch_words, st);
		if (e != cudaSuccess) return e;
		ws.scratch_zero = true;
	}
	if (n_rows > ws.row_cap) {
		if (ws.row_count) cudaFree(ws.row_count);
		if (ws.row_marks) cudaFree(ws.row_marks);
		ws.row_count = nullptr; ws.row_marks = nullptr; ws.row_cap = 0;
		e = cudaMalloc(&ws.row_count, n_rows * sizeof(unsigned int));
		if (e == cudaSuccess) e = cudaMalloc(&ws.row_marks, n_rows * 8 * sizeof(unsigned short));
		if (e != cudaSuccess) return e;
		ws.row_cap = n_rows;
		ws.rows_dirty = true;
	}
	if (ws.rows_dirty) {
		// between calls every counter is 0 and every slot holds 0xffff (= -1: a mark that covers nothing); the fill restores that
		e = cudaMemsetAsync(ws.row_count, 0, ws.row_cap * sizeof(unsigned int), st);
		if (e == cudaSuccess) e = cudaMemsetAsync(ws.row_marks, 0xff, ws.row_cap * 8 * sizeof(unsigned short), st);
		if (e != cudaSuccess) return e;
		ws.rows_dirty = false;
	}
	return cudaSuccess;
}
cudaError_t ensure_scratch(Workspace& ws, size_t words) {
	if (words <= ws.scratch_words) return cudaSuccess;
	if (ws.scratch) cudaFree(ws.scratch);
	ws.scratch = nullptr; ws.scratch_words = 0; ws.scratch_zero = false;
	cudaError_t e = cudaMalloc(&ws.scratch, words * sizeof(unsigned int));
	if (e == cudaSuccess) ws.scratch_words = words;
	return e;
}
}  // namespace voxb

// =================================================================================================
extern "C" {

const char* voxb200_last_error(void) { return g_err; }
const char* voxb200_version(void) { return "voxb200 0.1 (sm_100a)"; }

int voxb200_device_count(int* count) {
	if (!count) return fail(VOXB200_EINVAL, "count is NULL");
	*count = 0;
	cudaError_t e = cudaGetDeviceCount(count);
	if (e != cudaSuccess) { *count = 0; return fail_cuda(e, "cudaGetDeviceCount"); }
	return VOXB200_OK;
}

int voxb200_init(int device) {
	int n = 0;
	int rc = voxb200_device_count(&n);
	if (rc) return rc;
	if (n < 1) return fail(VOXB200_ENODEVICE, "no CUDA device found");
	if (device < 0 || device >= n) return fail(VOXB200_EINVAL, "device %d out of range (have %d)", device, n);
	CU(cudaSetDevice(device));
	Workspace* ws;
	return current_ws(&ws);
}

// util.h:80-110 (createMeshBBCube) then util.h:56-61 (voxinfo ctor): host binary32, no contraction
// (this TU's host side is built with -ffp-contract=off and no -march).
int voxb200_make_grid(const float mesh_min[3], const float mesh_max[3], unsigned int gridsize, size_t n_triangles, voxb200_grid* out) {
	if (!mesh_min || !mesh_max || !out) return fail(VOXB200_EINVAL, "NULL argument");
	if (gridsize == 0) return fail(VOXB200_EINVAL, "gridsize is 0");
	float len[3];
	for (int k = 0; k < 3; k++) len[k] = mesh_max[k] - mesh_min[k];
	const float longest = std::max(len[0], std::max(len[1], len[2]));
	float lo[3], hi[3];
	for (int k = 0; k < 3; k++) {
		lo[k] = mesh_min[k];
		hi[k] = mesh_max[k];
		if (longest != len[k]) {                         // pad the short axes symmetrically to the cube
			const float delta = longest - len[k];
			const float half = delta / 2.0f;
			lo[k] = mesh_min[k] - half;
			hi[k] = mesh_max[k] + half;
		}
	}
	for (int k = 0; k < 3; k++) {                        // the 1/10001 pad that keeps geometry off voxel faces
		const float side = hi[k] - lo[k];
		const float eps = side / 10001.0f;
		lo[k] = lo[k] - eps;
		hi[k] = hi[k] + eps;
	}
	memset(out, 0, sizeof(*out));
	for (int k = 0; k < 3; k++) {
		out->bbox_min[k] = lo[k];
		out->bbox_max[k] = hi[k];
		out->gridsize[k] = gridsize;
		const float side = hi[k] - lo[k];
		out->unit[k] = side / (float)gridsize;
	}
	out->n_triangles = n_triangles;
	return VOXB200_OK;
}

// main.cpp:190 — ceil(G^3 / 32.0f) * 4 with the reference's float division.  Above 2^24 voxels the conversion of G^3 to
// binary32 can round DOWN, and the reference's table then holds fewer than G^3 bits (629 of the grid sizes 1..2048, none
// of them a multiple of 32: G = 257, 513, 1025 are one bit short, G = 1026 eight): its far-corner voxels index the word one
// past the allocation.  A size API must not hand out a table the kernels can overrun, so this returns the larger of the
// reference's value and the exact ceil(G^3 / 32) * 4; voxb200_reference_table_bytes() keeps the reference's own number
// for callers (the C++ drop-in symbols) whose buffer was sized by the reference's main().
size_t voxb200_reference_table_bytes(unsigned int gridsize) {
	const size_t g = gridsize;
	return static_cast<size_t>(ceil((g * g * g) / 32.0f) * 4);
}
size_t voxb200_table_bytes(unsigned int gridsize) {
	const size_t g = gridsize;
	const size_t exact = ((g * g * g + 31) / 32) * 4;
	const size_t ref = voxb200_reference_table_bytes(gridsize);
	return exact > ref ? exact : ref;
}

uint64_t voxb200_morton_encode(unsigned int x, unsigned int y, unsigned int z) { return morton3(x, y, z); }

int voxb200_partition(unsigned int G, int morton, int part, int n_parts, voxb200_region* out, size_t* region_bytes) {
	if (!out) return fail(VOXB200_EINVAL, "out is NULL");
	if (G == 0 || n_parts < 1 || part < 0 || part >= n_parts) return fail(VOXB200_EINVAL, "bad partition %d of %d", part, n_parts);
	if (n_parts == 1) {
		for (int k = 0; k < 3; k++) { out->lo[k] = 0; out->hi[k] = (int)G; }
		if (region_bytes) *region_bytes = voxb200_table_bytes(G);
		return VOXB200_OK;
	}
	if (!morton) {
		if ((unsigned)n_parts > G) return fail(VOXB200_EINVAL, "more slabs (%d) than z-slices (%u)", n_parts, G);
		if (((unsigned long long)G * G) & 31ull) return fail(VOXB200_EINVAL, "z-slabs of a %u^3 grid are not word aligned", G);
		out->lo[0] = out->lo[1] = 0; out->hi[0] = out->hi[1] = (int)G;
		out->lo[2] = (int)(((unsigned long long)G * part) / n_parts);
		out->hi[2] = (int)(((unsigned long long)G * (part + 1)) / n_parts);
		if (region_bytes) *region_bytes = (size_t)G * G * (size_t)(out->hi[2] - out->lo[2]) / 8;
		return VOXB200_OK;
	}
	if ((G & (G - 1)) || (n_parts & (n_parts - 1))) return fail(VOXB200_EINVAL, "morton partition needs power-of-two grid and part count");
	const int m = ilog2(G), k = 3 * m - ilog2((unsigned)n_parts);      // k free low bits per part
	if (k < 5) return fail(VOXB200_EINVAL, "too many parts for a %u^3 morton table", G);
	const int a = (k + 2) / 3, b = (k + 1) / 3, c = k / 3;             // x gets bits 0,3,..; y 1,4,..; z 2,5,..
	const unsigned long long first = (unsigned long long)part << k;
	out->lo[0] = (int)compact3(first); out->lo[1] = (int)compact3(first >> 1); out->lo[2] = (int)compact3(first >> 2);
	out->hi[0] = out->lo[0] + (1 << a); out->hi[1] = out->lo[1] + (1 << b); out->hi[2] = out->lo[2] + (1 << c);
	if (region_bytes) *region_bytes = (size_t)1 << (k - 3);
	return VOXB200_OK;
}

int voxb200_malloc(void** dptr, size_t bytes) {
	if (!dptr) return fail(VOXB200_EINVAL, "dptr is NULL");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	CU(cudaMalloc(dptr, bytes ? bytes : 1));
	return VOXB200_OK;
}
int voxb200_free(void* dptr) {
	if (dptr) CU(cudaFree(dptr));
	return VOXB200_OK;
}
int voxb200_memcpy_d2h(void* host, const void* dptr, size_t bytes, void* stream) {
	if (!host || !dptr) return fail(VOXB200_EINVAL, "NULL pointer");
	CU(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	CU(cudaStreamSynchronize((cudaStream_t)stream));
	return VOXB200_OK;
}

int voxb200_upload_soup(const float* host_tris9, size_t n_triangles, int soa4, float** d_tris, void* stream) {
	if (!d_tris || (!host_tris9 && n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = ensure_host_path(hp);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const size_t soup_bytes = n_triangles * 9 * sizeof(float);
	float* d_soup = nullptr;
	CU(cudaMalloc(&d_soup, soup_bytes ? soup_bytes : 16));
	rc = h2d(hp, d_soup, host_tris9, soup_bytes, st);
	if (rc) { cudaFree(d_soup); return rc; }
	if (!soa4) { *d_tris = d_soup; return VOXB200_OK; }
	float* d_soa = nullptr;
	cudaError_t e = cudaMalloc(&d_soa, n_triangles ? n_triangles * 12 * sizeof(float) : 16);
	if (e != cudaSuccess) { cudaFree(d_soup); return fail_cuda(e, "cudaMalloc(soa4)"); }
	e = launch_soup_to_soa4(d_soup, d_soa, n_triangles, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(d_soup);
	if (e != cudaSuccess) { cudaFree(d_soa); return fail_cuda(e, "soup -> soa4"); }
	*d_tris = d_soa;
	return VOXB200_OK;
}

int voxb200_upload_indexed(const float* host_verts, size_t n_verts, const int32_t* host_faces, size_t n_faces,
                           int soa4, float** d_tris, float mesh_min[3], float mesh_max[3], void* stream) {
	if (!d_tris || !host_verts || (!host_faces && n_faces) || n_verts == 0) return fail(VOXB200_EINVAL, "NULL / empty mesh");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = ensure_host_path(hp);
	if (rc) return rc;
	// (face indices are range-checked on the device while they are expanded: counted, clamped, reported below)
	cudaStream_t st = (cudaStream_t)stream;
	float* d_verts = nullptr; int* d_faces = nullptr; float* d_out = nullptr; float* d_mm = nullptr;
	unsigned long long bad_faces = 0;
	cudaError_t e = cudaMalloc(&d_verts, n_verts * 3 * sizeof(float));
	if (e == cudaSuccess) e = cudaMalloc(&d_faces, n_faces ? n_faces * 3 * sizeof(int) : 16);
	if (e == cudaSuccess) e = cudaMalloc(&d_out, n_faces ? n_faces * (soa4 ? 12 : 9) * sizeof(float) : 16);
	if (e == cudaSuccess) e = cudaMalloc(&d_mm, 8 * sizeof(float));          // 6 floats of bbox + the bad-face counter
	if (e == cudaSuccess) {
		unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(d_mm + 6);
		e = cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st);
		rc = h2d(hp, d_verts, host_verts, n_verts * 3 * sizeof(float), st);
		if (!rc) rc = h2d(hp, d_faces, host_faces, n_faces * 3 * sizeof(int), st);
		if (!rc && e == cudaSuccess) {
			e = launch_expand_indexed(d_verts, d_faces, n_faces, n_verts, soa4 != 0, d_out, st, d_bad);
			if (e == cudaSuccess) e = cudaMemcpyAsync(&bad_faces, d_bad, sizeof(bad_faces), cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess && mesh_min && mesh_max) {
				e = launch_bbox_reduce(d_verts, n_verts, d_mm, st);
				float mm[6];
				if (e == cudaSuccess) e = cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st);
				if (e == cudaSuccess) e = cudaStreamSynchronize(st);
				if (e == cudaSuccess) for (int k = 0; k < 3; k++) { mesh_min[k] = mm[k]; mesh_max[k] = mm[3 + k]; }
			}
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		}
	}
	cudaFree(d_verts); cudaFree(d_faces); cudaFree(d_mm);
	if (rc) { cudaFree(d_out); return rc; }
	if (e != cudaSuccess) { cudaFree(d_out); return fail_cuda(e, "voxb200_upload_indexed"); }
	if (bad_faces) { cudaFree(d_out); return fail(VOXB200_EINVAL, "%llu faces have a vertex index out of range [0, %zu)", bad_faces, n_verts); }
	*d_tris = d_out;
	return VOXB200_OK;
}

int voxb200_route_triangles(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* region,
                            float** d_routed, size_t* n_routed, void* stream) {
	if (!grid || !d_routed || !n_routed || (!d_tris9 && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	if (flags & VOXB200_TRIS_SOA4) return fail(VOXB200_EINVAL, "routing takes the 9-float soup");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	GridParams g;
	size_t region_words = 0;
	rc = resolve_region(grid, region, (flags & VOXB200_MORTON) != 0, &g, &region_words);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	float* out = nullptr;
	CU(cudaMalloc(&out, grid->n_triangles ? grid->n_triangles * 9 * sizeof(float) : 16));
	cudaError_t e = launch_route(g, (flags & VOXB200_SOLID) != 0, d_tris9, out, ws->counters + kCtrClaim, st);
	unsigned long long n = 0;
	if (e == cudaSuccess) e = cudaMemcpyAsync(&n, ws->counters + kCtrClaim, sizeof(n), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) { cudaFree(out); return fail_cuda(e, "voxb200_route_triangles"); }
	*d_routed = out;
	*n_routed = (size_t)n;
	return VOXB200_OK;
}

int voxb200_sort_triangles(const voxb200_grid* grid, const float* d_tris9, float** d_sorted9, void* stream) {
	if (!grid || !d_sorted9 || (!d_tris9 && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	if (grid->n_triangles > 0xfffffffeull) return fail(VOXB200_EINVAL, "more than 2^32 triangles");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	GridParams g;
	size_t region_words = 0;
	rc = resolve_region(grid, nullptr, false, &g, &region_words);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	float* out = nullptr;
	unsigned int *keys = nullptr, *hist = nullptr;
	cudaError_t e = cudaMalloc(&out, grid->n_triangles ? grid->n_triangles * 9 * sizeof(float) : 16);
	if (e == cudaSuccess) e = cudaMalloc(&keys, grid->n_triangles ? grid->n_triangles * sizeof(unsigned int) : 16);
	if (e == cudaSuccess) e = cudaMalloc(&hist, (size_t)g.G * sizeof(unsigned int));
	if (e == cudaSuccess) e = launch_layer_sort(g, d_tris9, out, keys, hist, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);            // the scratch below is freed on return
	if (keys) cudaFree(keys);
	if (hist) cudaFree(hist);
	if (e != cudaSuccess) { if (out) cudaFree(out); return fail_cuda(e, "voxb200_sort_triangles"); }
	*d_sorted9 = out;
	return VOXB200_OK;
}

int voxb200_route_triangles_multi(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* regions,
                                  int n_regions, float* d_out, size_t out_capacity, size_t* counts, void* stream) {
	if (!grid || !regions || !counts || (!d_tris9 && grid->n_triangles) || (!d_out && out_capacity)) return fail(VOXB200_EINVAL, "NULL pointer");
	if (n_regions < 1 || n_regions > 32) return fail(VOXB200_EINVAL, "1..32 regions per call (got %d)", n_regions);
	if (flags & VOXB200_TRIS_SOA4) return fail(VOXB200_EINVAL, "routing takes the 9-float soup");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	GridParams g;
	size_t words = 0;
	int lo[32][3], hi[32][3];
	for (int r = 0; r < n_regions; r++) {
		rc = resolve_region(grid, &regions[r], (flags & VOXB200_MORTON) != 0, &g, &words);      // validates each region
		if (rc) return rc;
		for (int k = 0; k < 3; k++) { lo[r][k] = regions[r].lo[k]; hi[r][k] = regions[r].hi[k]; }
	}
	if (grid->n_triangles > ws->route_cap) {
		if (ws->route_masks) cudaFree(ws->route_masks);
		ws->route_masks = nullptr; ws->route_cap = 0;
		CU(cudaMalloc(&ws->route_masks, grid->n_triangles * sizeof(unsigned int)));
		ws->route_cap = grid->n_triangles;
	}
	if (!ws->route_counts) CU(cudaMalloc(&ws->route_counts, 64 * sizeof(unsigned long long)));
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e = launch_route_count(g, (flags & VOXB200_SOLID) != 0, lo, hi, n_regions, d_tris9, ws->route_masks, ws->route_counts, st);
	unsigned long long c[32];
	if (e == cudaSuccess) e = cudaMemcpyAsync(c, ws->route_counts, sizeof(c), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) return fail_cuda(e, "voxb200_route_triangles_multi (count)");
	unsigned long long cursors[32], total = 0;
	for (int r = 0; r < 32; r++) { cursors[r] = total; if (r < n_regions) { counts[r] = (size_t)c[r]; total += c[r]; } }
	if (total > out_capacity) return fail(VOXB200_EINVAL, "routed triangles (%llu) exceed the output capacity (%zu)", total, out_capacity);
	e = cudaMemcpyAsync(ws->route_counts + 32, cursors, sizeof(cursors), cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) e = launch_route_scatter(grid->n_triangles, n_regions, d_tris9, ws->route_masks, d_out, ws->route_counts + 32, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // `cursors` lives on this stack frame
	if (e != cudaSuccess) return fail_cuda(e, "voxb200_route_triangles_multi (scatter)");
	return VOXB200_OK;
}

int voxb200_surface(const voxb200_grid* grid, const float* d_tris, unsigned int* d_table, unsigned int flags,
                    const voxb200_region* region, void* stream) {
	return run_path(false, grid, d_tris, d_table, flags, region, (cudaStream_t)stream);
}
int voxb200_solid(const voxb200_grid* grid, const float* d_tris, unsigned int* d_table, unsigned int flags,
                  const voxb200_region* region, void* stream) {
	return run_path(true, grid, d_tris, d_table, flags, region, (cudaStream_t)stream);
}

// timing of the host entry points: [0] upload, [1] voxelization (device events); [3] the whole call on the host's clock; [2] the
// rest = the table's way back (dense copy, or compaction + pairs + the host threads' expansion: readback.cu)
static int host_timing(HostPath& hp, std::chrono::steady_clock::time_point t_call, float timing_ms[4]) {
	if (!timing_ms) return VOXB200_OK;
	CU(cudaEventElapsedTime(&timing_ms[0], hp.ev[0], hp.ev[1]));
	CU(cudaEventElapsedTime(&timing_ms[1], hp.ev[1], hp.ev[2]));
	timing_ms[3] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_call).count();
	timing_ms[2] = timing_ms[3] - timing_ms[0] - timing_ms[1];
	return VOXB200_OK;
}

int voxb200_voxelize_host(const voxb200_grid* grid, const float* host_tris9, unsigned int* host_table,
                          unsigned int flags, const voxb200_region* region, float timing_ms[4]) {
	const auto t_call = std::chrono::steady_clock::now();
	if (!grid || !host_table || (!host_tris9 && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	if (flags & VOXB200_TRIS_SOA4) return fail(VOXB200_EINVAL, "voxb200_voxelize_host takes the 9-float soup");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = ensure_host_path(hp);
	if (rc) return rc;
	GridParams g;
	size_t region_words = 0;
	rc = resolve_region(grid, region, (flags & VOXB200_MORTON) != 0, &g, &region_words);
	if (rc) return rc;
	const size_t tris_bytes = (size_t)grid->n_triangles * 9 * sizeof(float);
	const size_t table_bytes = region_words * sizeof(unsigned int);
	if (tris_bytes > hp.tris_bytes) {
		if (hp.d_tris) cudaFree(hp.d_tris);
		hp.d_tris = nullptr; hp.tris_bytes = 0;
		CU(cudaMalloc(&hp.d_tris, tris_bytes));
		hp.tris_bytes = tris_bytes;
	}
	if (table_bytes > hp.table_bytes) {
		if (hp.d_table) cudaFree(hp.d_table);
		hp.d_table = nullptr; hp.table_bytes = 0;
		CU(cudaMalloc(&hp.d_table, table_bytes));
		hp.table_bytes = table_bytes;
	}
	cudaStream_t st = hp.stream;
	ReadbackGuard guard{hp.rb};
	CU(cudaEventRecord(hp.ev[0], st));
	rc = h2d(hp, hp.d_tris, host_tris9, tris_bytes, st);
	if (rc) return rc;
	if (!(flags & VOXB200_SOLID)) readback_prezero(hp.rb, host_table, region_words, 0, st, true);      // surface tables are sparse: zero-fill the host table meanwhile
	CU(cudaEventRecord(hp.ev[1], st));
	const unsigned int path_flags = flags & (VOXB200_MORTON);
	rc = run_path((flags & VOXB200_SOLID) != 0, grid, hp.d_tris, hp.d_table, path_flags, region, st);
	if (rc) return rc;
	CU(cudaEventRecord(hp.ev[2], st));
	unsigned long long overflow = 0ull;
	CU(cudaMemcpyAsync(&overflow, ws->counters + kCtrQueueOverflow, sizeof(overflow), cudaMemcpyDeviceToHost, st));
	rc = readback_table(hp.rb, hp.d_table, region_words, host_table, st, 0);
	if (rc) return rc;
	if (overflow) return fail(VOXB200_EINVAL, "the mesh queues more than 2^32 (y,z) rows / sample blocks for the large-triangle path at this grid size: table contents undefined");
	return host_timing(hp, t_call, timing_ms);
}

// Upload + prepare + voxelize of the host entry points that take an indexed mesh; the table is left in hp.d_table (region_words
// words), events 0..2 are recorded.  host_table: where the dense table will go (nullptr when the caller wants the non-zero words).
static int host_indexed_to_device(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                  unsigned int* host_table, unsigned int flags, const voxb200_region* region, Workspace* ws, HostPath& hp,
                                  size_t* region_words_out) {
	GridParams g;
	size_t region_words = 0;
	int rc = resolve_region(grid, region, (flags & VOXB200_MORTON) != 0, &g, &region_words);
	if (rc) return rc;
	*region_words_out = region_words;
	const size_t n_faces = grid->n_triangles;
	const size_t table_bytes = region_words * sizeof(unsigned int);
	// Surface, linear order, tileable grid: the upload path bins the faces straight into the tile records of a prepared mesh (no
	// intermediate soup) and the tile-owner kernel writes the table; everything else expands to a soup and runs the one-shot kernels.
	const bool tiles = mesh_tileable(g, flags & (VOXB200_SOLID | VOXB200_MORTON)) && n_faces > 0;
	if ((rc = grow(&hp.d_verts, &hp.verts_bytes, n_verts * 3 * sizeof(float)))) return rc;
	if ((rc = grow(&hp.d_faces, &hp.faces_bytes, n_faces * 3 * sizeof(int) + 16))) return rc;
	if (!tiles && (rc = grow(&hp.d_tris, &hp.tris_bytes, n_faces * 9 * sizeof(float) + 16))) return rc;
	if ((rc = grow(&hp.d_table, &hp.table_bytes, table_bytes))) return rc;
	cudaStream_t st = hp.stream;
	CU(cudaEventRecord(hp.ev[0], st));
	rc = h2d(hp, hp.d_verts, host_verts, n_verts * 3 * sizeof(float), st);
	if (!rc) rc = h2d(hp, hp.d_faces, host_faces, n_faces * 3 * sizeof(int), st);
	if (rc) return rc;
	if (host_table && !(flags & VOXB200_SOLID)) readback_prezero(hp.rb, host_table, region_words, 0, st, true);      // surface tables are sparse: zero-fill the host table meanwhile
	if (tiles) {
		const bool same = hp.mesh && memcmp(&hp.mesh_grid, grid, sizeof(*grid)) == 0 && hp.mesh_has_region == (region != nullptr) &&
		                  (!region || memcmp(&hp.mesh_region, region, sizeof(*region)) == 0);
		if (same) {
			rc = voxb200_mesh_update_indexed(hp.mesh, hp.d_verts, n_verts, hp.d_faces, st);
		} else {
			if (hp.mesh) { voxb200_mesh_destroy(hp.mesh); hp.mesh = nullptr; }
			rc = voxb200_mesh_create_indexed(grid, hp.d_verts, n_verts, hp.d_faces, 0u, region, &hp.mesh, st);
			if (!rc) { hp.mesh_grid = *grid; hp.mesh_has_region = region != nullptr; if (region) hp.mesh_region = *region; }
		}
		if (rc) return rc;
		CU(cudaEventRecord(hp.ev[1], st));
		rc = voxb200_mesh_voxelize(hp.mesh, hp.d_table, 0u, st);
		if (rc) return rc;
	} else {
		if (!hp.d_bad) CU(cudaMalloc(&hp.d_bad, sizeof(unsigned long long)));
		CU(cudaMemsetAsync(hp.d_bad, 0, sizeof(unsigned long long), st));
		cudaError_t e = launch_expand_indexed(hp.d_verts, hp.d_faces, n_faces, n_verts, false, hp.d_tris, st, hp.d_bad);
		if (e != cudaSuccess) return fail_cuda(e, "expand_indexed");
		CU(cudaEventRecord(hp.ev[1], st));
		rc = run_path((flags & VOXB200_SOLID) != 0, grid, hp.d_tris, hp.d_table, flags & VOXB200_MORTON, region, st);
		if (rc) return rc;
	}
	CU(cudaEventRecord(hp.ev[2], st));
	hp.last_tiles = tiles;
	return VOXB200_OK;
}
// after the read-back has synchronised the stream: did the large-triangle queue overflow?
static int host_indexed_check(Workspace* ws, HostPath& hp) {
	unsigned long long overflow = 0ull;
	if (hp.last_tiles) {
		uint64_t c[4];
		const int rc = voxb200_mesh_counters(hp.mesh, c);
		if (rc) return rc;
		overflow = c[1] == ~0ull;
	} else {
		CU(cudaMemcpy(&overflow, ws->counters + kCtrQueueOverflow, sizeof(overflow), cudaMemcpyDeviceToHost));
		unsigned long long bad_faces = 0ull;
		if (hp.d_bad) CU(cudaMemcpy(&bad_faces, hp.d_bad, sizeof(bad_faces), cudaMemcpyDeviceToHost));
		if (bad_faces) return fail(VOXB200_EINVAL, "%llu faces have a vertex index out of range: table contents undefined", bad_faces);
	}
	if (overflow) return fail(VOXB200_EINVAL, "the mesh queues more than 2^32 (y,z) rows / sample blocks for the large-triangle path at this grid size: table contents undefined");
	return VOXB200_OK;
}

int voxb200_voxelize_host_indexed(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                  unsigned int* host_table, unsigned int flags, const voxb200_region* region, float timing_ms[4]) {
	const auto t_call = std::chrono::steady_clock::now();
	if (!grid || !host_table || !host_verts || (!host_faces && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = ensure_host_path(hp);
	if (rc) return rc;
	ReadbackGuard guard{hp.rb};
	size_t region_words = 0;
	rc = host_indexed_to_device(grid, host_verts, n_verts, host_faces, host_table, flags, region, ws, hp, &region_words);
	if (rc) return rc;
	rc = readback_table(hp.rb, hp.d_table, region_words, host_table, hp.stream, 0);
	if (rc) return rc;
	if ((rc = host_indexed_check(ws, hp))) return rc;
	return host_timing(hp, t_call, timing_ms);
}

int voxb200_voxelize_host_nonzero(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                  unsigned int flags, const voxb200_region* region, const voxb200_word** words, size_t* n_words, float timing_ms[4]) {
	const auto t_call = std::chrono::steady_clock::now();
	if (!grid || !words || !n_words || !host_verts || (!host_faces && grid->n_triangles)) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = ensure_host_path(hp);
	if (rc) return rc;
	size_t region_words = 0;
	rc = host_indexed_to_device(grid, host_verts, n_verts, host_faces, nullptr, flags, region, ws, hp, &region_words);
	if (rc) return rc;
	const void* pairs = nullptr;
	rc = readback_pairs(hp.rb, hp.d_table, region_words, hp.stream, &pairs, n_words);
	if (rc) return rc;
	*words = static_cast<const voxb200_word*>(pairs);
	if ((rc = host_indexed_check(ws, hp))) return rc;
	return host_timing(hp, t_call, timing_ms);
}

int voxb200_last_readback(uint64_t info[2]) {
	if (!info) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	info[0] = (uint64_t)g_hp[ws->device].rb.last_mode;
	info[1] = g_hp[ws->device].rb.last_nonzero;
	return VOXB200_OK;
}

int voxb200_download_table(const unsigned int* d_table, size_t table_words, unsigned int* host_table, void* stream, uint64_t info[2]) {
	if (!d_table || !host_table) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	HostPath& hp = g_hp[ws->device];
	rc = readback_table(hp.rb, d_table, table_words, host_table, (cudaStream_t)stream, 0);
	if (rc) return rc;
	if (info) { info[0] = (uint64_t)hp.rb.last_mode; info[1] = hp.rb.last_nonzero; }
	return VOXB200_OK;
}

int voxb200_extract_voxels(const unsigned int* d_table, size_t table_words, uint64_t first_voxel, uint64_t** d_indices, size_t* count, void* stream) {
	if (!d_table || !d_indices || !count) return fail(VOXB200_EINVAL, "NULL pointer");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const size_t blocks = extract_blocks(table_words);
	unsigned int* d_counts = nullptr;
	unsigned long long* d_offsets = nullptr;
	unsigned long long* d_out = nullptr;
	cudaError_t e = cudaMalloc(&d_counts, (blocks + 1) * sizeof(unsigned int));
	if (e == cudaSuccess) e = cudaMalloc(&d_offsets, (blocks + 1) * sizeof(unsigned long long));
	unsigned long long total = 0;
	if (e == cudaSuccess) e = launch_extract_count(d_table, table_words, d_counts, d_offsets, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_offsets + blocks, sizeof(total), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e == cudaSuccess) e = cudaMalloc(&d_out, (total ? total : 1) * sizeof(unsigned long long));
	if (e == cudaSuccess) e = launch_extract_write(d_table, table_words, d_offsets, first_voxel, d_out, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(d_counts);
	cudaFree(d_offsets);
	if (e != cudaSuccess) { cudaFree(d_out); return fail_cuda(e, "voxb200_extract_voxels"); }
	*d_indices = reinterpret_cast<uint64_t*>(d_out);
	*count = (size_t)total;
	return VOXB200_OK;
}

int voxb200_release(void) {
	int dev = -1;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return VOXB200_OK; }
	Workspace& ws = g_ws[dev];
	HostPath& hp = g_hp[dev];
	if (ws.device == dev) CU(cudaDeviceSynchronize());
	if (hp.mesh) { voxb200_mesh_destroy(hp.mesh); hp.mesh = nullptr; }
	readback_free(hp.rb);
	multi_release_device(dev);
	void* dev_ptrs[] = {ws.counters, ws.queue, ws.setups, ws.dir, ws.route_masks, ws.route_counts, ws.scratch, ws.row_count, ws.row_marks,
	                    hp.d_tris, hp.d_table, hp.d_verts, hp.d_faces, hp.d_bad};
	for (void* p : dev_ptrs) if (p) cudaFree(p);
	for (void* p : hp.pinned) if (p) cudaFreeHost(p);
	if (ws.prof_ev) {
		for (int i = 0; i < kProfRing; i++) for (int k = 0; k < kProfEvents; k++) cudaEventDestroy(ws.prof_ev[i][k]);
		delete[] ws.prof_ev;
	}
	if (hp.stream) cudaStreamDestroy(hp.stream);
	for (auto& e : hp.ev) if (e) cudaEventDestroy(e);
	for (auto& e : hp.buf_free) if (e) cudaEventDestroy(e);
	ws = Workspace();
	hp = HostPath();
	cudaGetLastError();
	return VOXB200_OK;
}

uint64_t voxb200_launch_count(int reset) {
	const uint64_t v = g_launch_count;
	if (reset) g_launch_count = 0;
	return v;
}

int voxb200_set_profiling(int on) {
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	if (on && !ws->prof_ev) {
		ws->prof_ev = new cudaEvent_t[kProfRing][kProfEvents];
		for (int i = 0; i < kProfRing; i++)
			for (int k = 0; k < kProfEvents; k++) CU(cudaEventCreate(&ws->prof_ev[i][k]));
	}
	ws->prof_on = on != 0;
	ws->prof_calls = 0;
	return VOXB200_OK;
}

int voxb200_phase_ms(unsigned int call_index, float out[4]) {
	if (!out) return fail(VOXB200_EINVAL, "out is NULL");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	if (!ws->prof_ev || call_index >= ws->prof_calls || ws->prof_calls - call_index > (unsigned)kProfRing)
		return fail(VOXB200_EINVAL, "call %u is not in the profiling ring (%u calls recorded, ring of %d)", call_index, ws->prof_calls, kProfRing);
	cudaEvent_t* ev = ws->prof_ev[call_index % kProfRing];
	for (int k = 0; k < 4; k++) CU(cudaEventElapsedTime(&out[k], ev[k], ev[k + 1]));
	return VOXB200_OK;
}

int voxb200_last_counters(uint64_t out[4]) {
	if (!out) return fail(VOXB200_EINVAL, "out is NULL");
	Workspace* ws;
	int rc = current_ws(&ws);
	if (rc) return rc;
	unsigned long long c[kNumCounters];
	CU(cudaMemcpy(c, ws->counters, sizeof(c), cudaMemcpyDeviceToHost));
	out[0] = c[kCtrQueue] >> 32;
	out[1] = c[kCtrQueueOverflow] ? ~0ull : (c[kCtrQueue] & 0xffffffffull);
	out[2] = c[kCtrSolidClamp];
	out[3] = ws->last_row_lists ? 1 : 0;
	return VOXB200_OK;
}

}  // extern "C"
