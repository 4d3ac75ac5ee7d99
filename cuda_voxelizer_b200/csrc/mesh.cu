// mesh.cu — prepared meshes: the resident / per-frame interface of the hot path (the reference pitches
// per-frame voxelization, README.md:74, but offers only the one-shot voxelize(), voxelize.cu:192).
//
// A voxb200_mesh owns everything a voxelization of ONE mesh on ONE grid (and region) needs: the triangles re-ordered
// for the kernels, the plan of the tile-owner surface path (tiles.cu), the side soup of triangles that take the
// row-solver path, and a PRIVATE workspace — nothing on its hot path touches the library's per-device state, so
// handles are re-entrant: different handles may voxelize concurrently on different streams.
//
// Two schedules, chosen at creation:
//   TILES   surface, linear order, G a multiple of 256 (of 1024 above 1024), a region of at most 2^32 words and 2^20 tiles (all of a
//           4096^3 grid; 8192^3 in z-slabs of up to 2048 layers), z-range on tile boundaries: surface_tile_kernel
//           writes every table byte once + launch_surface(ACCUMULATE) over the side soup.
//   DIRECT  everything else (solid, morton, other grid sizes): the one-shot kernels on the handle's own copy of the
//           soup (z-layer ordered for the surface path), with the handle's own workspace.
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/voxb200.h"
#include "vox_internal.h"

using namespace voxb;

struct voxb200_mesh {
	int device = -1;
	voxb200_grid grid{};
	bool has_region = false;
	voxb200_region region{};
	unsigned int flags = 0;                 // VOXB200_SOLID | VOXB200_MORTON
	GridParams g{};
	size_t region_words = 0;
	bool tiles = false;
	Workspace ws;                           // private: queue, counters, solid scratch
	// TILES schedule
	TilePlan plan{};
	float* binned = nullptr; size_t binned_cap = 0;        // records (in floats)
	float* side = nullptr; size_t side_cap = 0;            // triangles
	size_t n_side = 0;
	unsigned int *cnt = nullptr, *off = nullptr, *order = nullptr, *empty = nullptr, *fill = nullptr;
	uint4* work = nullptr;
	size_t tiles_cap = 0;
	unsigned int* keys = nullptr; size_t keys_cap = 0;
	unsigned long long* totals = nullptr;
	unsigned long long host_totals[kPlanTotals] = {};
	// DIRECT schedule
	float* soup = nullptr; size_t soup_cap = 0;
	unsigned int *sort_keys = nullptr, *sort_hist = nullptr;
	uint64_t calls = 0;
};

namespace {

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return abi_fail_cuda(e_, #call); } while (0)

template <typename T>
int grow_dev(T** p, size_t* have, size_t want) {
	if (want <= *have && *p) return VOXB200_OK;
	if (*p) cudaFree(*p);
	*p = nullptr; *have = 0;
	CU(cudaMalloc(p, (want ? want : 1) * sizeof(T)));
	*have = want;
	return VOXB200_OK;
}

bool tileable(const voxb200_mesh& m) { return voxb::mesh_tileable(m.g, m.flags); }

}  // namespace

namespace voxb {
bool mesh_tileable(const GridParams& g, unsigned int flags) {
	if (flags & (VOXB200_SOLID | VOXB200_MORTON)) return false;
	// coordinates travel as 16 bits in the records, word offsets inside the region as 32 (8192^3 in slabs of at most 2048 layers),
	// tiles as 20 bits in the planning keys
	if (g.G < 256 || g.G > 32768 || !g.w32 || (g.G % (g.G < kTileXMax ? 256 : kTileXMax)) != 0 || (g.G < kTileXMax && (g.G & (g.G - 1)) != 0)) return false;
	if (g.rx0 != 0 || g.rx1 != g.G || g.ry0 != 0 || g.ry1 != g.G) return false;
	if ((g.rz0 % kTileZ) != 0 || (g.rz1 % kTileZ) != 0) return false;
	const int tile_x = g.G < kTileXMax ? g.G : kTileXMax;
	const unsigned long long tiles = (unsigned long long)(g.G / tile_x) * (unsigned long long)(g.G / kTileY) * (unsigned long long)((g.rz1 - g.rz0) / kTileZ);
	if (tiles > (1ull << 20)) return false;
	return true;
}
}  // namespace voxb

namespace {

// (Re)builds the handle's device state from a triangle source: a 9-float device soup, or device vertices + faces.
int prepare(voxb200_mesh& m, const float* d_soup, const float* d_verts, const int* d_faces, size_t n_verts, cudaStream_t st) {
	const GridParams& g = m.g;
	const size_t n = (size_t)g.n_tris;
	if (!m.tiles) {
		// DIRECT: own copy of the soup; the surface path gets it in z-layer order (same table, the atomics sweep the table)
		int rc = grow_dev(&m.soup, &m.soup_cap, n * 9 + 16);
		if (rc) return rc;
		const bool sort = !(m.flags & (VOXB200_SOLID | VOXB200_MORTON)) && n > 0;
		float* expanded = m.soup;
		float* tmp = nullptr;
		if (sort) { CU(cudaMalloc(&tmp, (n * 9 + 16) * sizeof(float))); expanded = tmp; }
		cudaError_t e = cudaSuccess;
		unsigned long long bad_faces = 0;
		if (d_faces) {
			if (!m.totals) CU(cudaMalloc(&m.totals, kPlanTotals * sizeof(unsigned long long)));
			e = cudaMemsetAsync(m.totals, 0, kPlanTotals * sizeof(unsigned long long), st);
			if (e == cudaSuccess) e = launch_expand_indexed(d_verts, d_faces, n, n_verts, false, expanded, st, m.totals + kPlanBadFaces);
			if (e == cudaSuccess) e = cudaMemcpyAsync(&bad_faces, m.totals + kPlanBadFaces, sizeof(bad_faces), cudaMemcpyDeviceToHost, st);
		}
		else if (n) e = cudaMemcpyAsync(expanded, d_soup, n * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st);
		if (e == cudaSuccess && sort) {
			size_t kc = m.keys_cap, hc = 0;
			int rc2 = grow_dev(&m.keys, &kc, n);
			m.keys_cap = kc;
			unsigned int* hist = nullptr;
			if (!rc2) rc2 = grow_dev(&hist, &hc, (size_t)g.G);
			if (rc2) { cudaFree(tmp); return rc2; }
			e = launch_layer_sort(g, tmp, m.soup, m.keys, hist, st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			cudaFree(hist);
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		if (tmp) cudaFree(tmp);
		if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh prepare (direct schedule)");
		if (bad_faces) return abi_fail(VOXB200_EINVAL, "%llu faces have a vertex index outside [0, %zu)", bad_faces, n_verts);
		return VOXB200_OK;
	}
	// TILES
	TileGeom tg;
	const int tile_x = g.G < kTileXMax ? g.G : kTileXMax;                   // 256, 512 or 1024: whole rows, or whole 128-byte lines
	tg.tx_shift = tile_x == 256 ? 8 : tile_x == 512 ? 9 : 10;
	tg.chunk_shift = tg.tx_shift - 12;
	for (int v = kTileY * kTileZ; v > 1; v >>= 1) tg.chunk_shift++;      // log2(tile_x * kTileY * kTileZ / 8 / 512)
	tg.G = g.G;
	tg.ntx = g.G / tile_x; tg.nty = g.G / kTileY; tg.ntz = (g.rz1 - g.rz0) / kTileZ; tg.tz0 = g.rz0 / kTileZ;
	tg.n_tiles = (unsigned int)tg.ntx * (unsigned int)tg.nty * (unsigned int)tg.ntz;
	tg.n_verts = d_faces ? (unsigned int)(n_verts > 0x7fffffffull ? 0x7fffffffull : n_verts) : 0u;
	const size_t nt = tg.n_tiles;
	if (nt + 1 > m.tiles_cap || !m.cnt) {
		for (unsigned int** p : {&m.cnt, &m.off, &m.order, &m.empty, &m.fill}) { if (*p) cudaFree(*p); *p = nullptr; }
		if (m.work) cudaFree(m.work);
		m.work = nullptr;
		m.tiles_cap = 0;
		for (unsigned int** p : {&m.cnt, &m.off, &m.order, &m.empty, &m.fill}) CU(cudaMalloc(p, (nt + 1) * sizeof(unsigned int)));
		CU(cudaMalloc(&m.work, (nt + 1) * sizeof(uint4)));
		m.tiles_cap = nt + 1;
	}
	if (!m.totals) CU(cudaMalloc(&m.totals, kPlanTotals * sizeof(unsigned long long)));
	int rc = grow_dev(&m.keys, &m.keys_cap, n);
	if (rc) return rc;
	// a tile whose run would keep one CTA busy for a large part of the whole kernel is not binned: its triangles take the side path
	const unsigned int cap = (unsigned int)(n / 256 > 8192 ? n / 256 : 8192);
	cudaError_t e = launch_tile_count(g, tg, d_soup, d_verts, d_faces, m.keys, m.cnt, m.totals, st);
	if (e == cudaSuccess) e = launch_tile_plan(tg, cap, m.cnt, m.off, m.order, m.work, m.empty, m.totals, m.fill, st);      // fill: free until the scatter clears it
	if (e == cudaSuccess) e = cudaMemcpyAsync(m.host_totals, m.totals, sizeof(m.host_totals), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh prepare (count / plan)");
	if (m.host_totals[kPlanBadFaces]) return abi_fail(VOXB200_EINVAL, "%llu faces have a vertex index outside [0, %zu)", m.host_totals[kPlanBadFaces], n_verts);
	const unsigned long long inst = m.host_totals[kPlanInstances];
	if (inst >= 0xfffffff0ull) return abi_fail(VOXB200_EINVAL, "more than 2^32 triangle instances in the tile plan");
	const size_t side_max = (size_t)(m.host_totals[kPlanBigDirect] + m.host_totals[kPlanHeavyInstances]);
	rc = grow_dev(&m.binned, &m.binned_cap, ((size_t)inst + 64) * 16);         // 64-byte records; padded: the last batch of a run reads 32 of them
	if (!rc) rc = grow_dev(&m.side, &m.side_cap, (side_max + 4) * 9);
	if (rc) return rc;
	e = launch_tile_scatter(g, tg, cap, d_soup, d_verts, d_faces, m.keys, m.cnt, m.off, m.fill, m.binned, m.side, m.totals, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(m.host_totals, m.totals, sizeof(m.host_totals), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh prepare (scatter)");
	m.n_side = (size_t)m.host_totals[kPlanSideFill];
	TilePlan& p = m.plan;
	p.geom = tg;
	p.soup = m.binned; p.cnt = m.cnt; p.off = m.off; p.order = m.order; p.work = m.work; p.empty = m.empty;
	p.n_work = (unsigned int)m.host_totals[kPlanWork];
	p.wide = m.host_totals[kPlanWide] != 0;
	p.n_empty = (unsigned int)m.host_totals[kPlanEmpty];
	p.zero_chunks = p.n_empty << tg.chunk_shift;
	const unsigned long long batches = m.host_totals[kPlanBatches];
	// every batch clears its share of the empty tiles; a mesh with few small triangles leaves the rest to zero-only blocks
	unsigned long long quota = batches ? ((unsigned long long)p.zero_chunks + batches - 1) / batches : 0;
	if (quota > 16) quota = 16;
	p.zero_quota = (unsigned int)quota;
	const unsigned long long covered = quota * batches;
	p.zero_rest_first = (unsigned int)(covered < p.zero_chunks ? covered : p.zero_chunks);
	p.n_zero_blocks = (p.zero_chunks - p.zero_rest_first + kZeroBlockChunks - 1) / kZeroBlockChunks;
	cudaError_t qe = ensure_queue(m.ws, m.n_side);
	if (qe != cudaSuccess) return abi_fail_cuda(qe, "voxb200_mesh prepare (queue)");
	return VOXB200_OK;
}

int create_common(const voxb200_grid* grid, unsigned int flags, const voxb200_region* region, voxb200_mesh** out) {
	if (!grid || !out) return abi_fail(VOXB200_EINVAL, "NULL grid / out pointer");
	if (flags & ~(VOXB200_SOLID | VOXB200_MORTON)) return abi_fail(VOXB200_EINVAL, "voxb200_mesh_create takes VOXB200_SOLID and VOXB200_MORTON only");
	Workspace* cur;
	int rc = abi_current_ws(&cur);          // also checks the device
	if (rc) return rc;
	voxb200_mesh* m = new (std::nothrow) voxb200_mesh();
	if (!m) return abi_fail(VOXB200_ENOMEM, "out of host memory");
	m->device = cur->device;
	m->grid = *grid;
	m->flags = flags;
	if (region) { m->has_region = true; m->region = *region; }
	rc = abi_resolve_region(grid, region, (flags & VOXB200_MORTON) != 0, &m->g, &m->region_words);
	if (!rc) rc = abi_init_workspace(m->ws, m->device);
	if (rc) { delete m; return rc; }
	m->tiles = tileable(*m);
	*out = m;
	return VOXB200_OK;
}

void destroy(voxb200_mesh* m) {
	for (void* p : {(void*)m->binned, (void*)m->side, (void*)m->cnt, (void*)m->off, (void*)m->order, (void*)m->work, (void*)m->empty,
	                (void*)m->fill, (void*)m->keys, (void*)m->totals, (void*)m->soup, (void*)m->sort_keys, (void*)m->sort_hist})
		if (p) cudaFree(p);
	abi_free_workspace(m->ws);
	delete m;
}

int check_device(const voxb200_mesh* m) {
	int dev = -1;
	CU(cudaGetDevice(&dev));
	if (dev != m->device) return abi_fail(VOXB200_EINVAL, "the mesh lives on device %d but device %d is current", m->device, dev);
	return VOXB200_OK;
}

}  // namespace

extern "C" {

int voxb200_mesh_create(const voxb200_grid* grid, const float* d_tris9, unsigned int flags, const voxb200_region* region,
                        voxb200_mesh** out, void* stream) {
	if (grid && !d_tris9 && grid->n_triangles) return abi_fail(VOXB200_EINVAL, "NULL triangle pointer");
	voxb200_mesh* m = nullptr;
	int rc = create_common(grid, flags, region, &m);
	if (rc) return rc;
	rc = prepare(*m, d_tris9, nullptr, nullptr, 0, (cudaStream_t)stream);
	if (rc) { destroy(m); return rc; }
	*out = m;
	return VOXB200_OK;
}

int voxb200_mesh_create_indexed(const voxb200_grid* grid, const float* d_verts, size_t n_verts, const int32_t* d_faces, unsigned int flags,
                                const voxb200_region* region, voxb200_mesh** out, void* stream) {
	if (grid && grid->n_triangles && (!d_verts || !d_faces || n_verts == 0)) return abi_fail(VOXB200_EINVAL, "NULL / empty indexed mesh");
	voxb200_mesh* m = nullptr;
	int rc = create_common(grid, flags, region, &m);
	if (rc) return rc;
	rc = prepare(*m, nullptr, d_verts, reinterpret_cast<const int*>(d_faces), n_verts, (cudaStream_t)stream);
	if (rc) { destroy(m); return rc; }
	*out = m;
	return VOXB200_OK;
}

int voxb200_mesh_update(voxb200_mesh* m, const float* d_tris9, void* stream) {
	if (!m || (!d_tris9 && m->g.n_tris)) return abi_fail(VOXB200_EINVAL, "NULL mesh / triangle pointer");
	int rc = check_device(m);
	if (rc) return rc;
	return prepare(*m, d_tris9, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int voxb200_mesh_update_indexed(voxb200_mesh* m, const float* d_verts, size_t n_verts, const int32_t* d_faces, void* stream) {
	if (!m || (m->g.n_tris && (!d_verts || !d_faces || n_verts == 0))) return abi_fail(VOXB200_EINVAL, "NULL mesh / vertex / face pointer");
	int rc = check_device(m);
	if (rc) return rc;
	return prepare(*m, nullptr, d_verts, reinterpret_cast<const int*>(d_faces), n_verts, (cudaStream_t)stream);
}

int voxb200_mesh_voxelize(voxb200_mesh* m, unsigned int* d_table, unsigned int flags, void* stream) {
	if (!m || !d_table) return abi_fail(VOXB200_EINVAL, "NULL mesh / table pointer");
	if (flags & ~VOXB200_ACCUMULATE) return abi_fail(VOXB200_EINVAL, "voxb200_mesh_voxelize takes VOXB200_ACCUMULATE only (order and mode were fixed at creation)");
	int rc = check_device(m);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	LaunchOpts o;
	o.morton = (m->flags & VOXB200_MORTON) != 0;
	o.accumulate = (flags & VOXB200_ACCUMULATE) != 0;
	o.soa4 = false;
	m->calls++;
	cudaError_t e;
	if (!m->tiles) {
		e = (m->flags & VOXB200_SOLID) ? launch_solid(m->ws, m->g, m->soup, d_table, m->region_words, o, st)
		                               : launch_surface(m->ws, m->g, m->soup, d_table, m->region_words, o, st);
		if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh_voxelize (direct schedule)");
		return VOXB200_OK;
	}
	if (reinterpret_cast<uintptr_t>(d_table) & 15u) return abi_fail(VOXB200_EINVAL, "the table must be 16-byte aligned");
	// optional per-phase timing through the library's profiling ring (bench.py): [0] -, [1] tile kernel, [2] side path, [3] -
	Workspace* prof = nullptr;
	if (abi_current_ws(&prof) == VOXB200_OK && prof->prof_on) { prof_mark(*prof, 0, st); prof_mark(*prof, 1, st); } else prof = nullptr;
	e = launch_surface_tiles(m->g, m->plan, d_table, o.accumulate, st);
	if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh_voxelize (tile kernel)");
	if (prof) prof_mark(*prof, 2, st);
	if (m->n_side) {
		GridParams gs = m->g;
		gs.n_tris = m->n_side;
		LaunchOpts os = o;
		os.accumulate = true;                 // the tile kernel has written every byte of the region; the side path ORs into it
		e = launch_surface(m->ws, gs, m->side, d_table, m->region_words, os, st);
		if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_mesh_voxelize (side path)");
	}
	if (prof) { prof_mark(*prof, 3, st); prof_mark(*prof, 4, st); prof->prof_calls++; }
	return VOXB200_OK;
}

int voxb200_mesh_info(const voxb200_mesh* m, uint64_t out[8]) {
	if (!m || !out) return abi_fail(VOXB200_EINVAL, "NULL pointer");
	memset(out, 0, 8 * sizeof(uint64_t));
	out[0] = m->tiles ? 1 : 0;
	if (m->tiles) {
		out[1] = m->plan.geom.n_tiles; out[2] = m->plan.n_work; out[3] = m->host_totals[kPlanInstances];
		out[4] = m->n_side; out[5] = m->host_totals[kPlanBatches]; out[6] = m->plan.zero_quota; out[7] = m->plan.n_zero_blocks;
	}
	return VOXB200_OK;
}

int voxb200_mesh_counters(const voxb200_mesh* m, uint64_t out[4]) {
	if (!m || !out) return abi_fail(VOXB200_EINVAL, "NULL pointer");
	int rc = check_device(m);
	if (rc) return rc;
	unsigned long long c[kNumCounters];
	CU(cudaMemcpy(c, m->ws.counters, sizeof(c), cudaMemcpyDeviceToHost));
	out[0] = c[kCtrQueue] >> 32;
	out[1] = c[kCtrQueueOverflow] ? ~0ull : (c[kCtrQueue] & 0xffffffffull);
	out[2] = c[kCtrSolidClamp];
	out[3] = m->ws.last_row_lists ? 1 : 0;
	return VOXB200_OK;
}

int voxb200_mesh_destroy(voxb200_mesh* m) {
	if (!m) return VOXB200_OK;
	int dev = -1;
	const int mine = m->device;
	cudaGetDevice(&dev);
	if (dev != mine) cudaSetDevice(mine);
	cudaDeviceSynchronize();
	destroy(m);
	if (dev >= 0 && dev != mine) cudaSetDevice(dev);
	cudaGetLastError();
	return VOXB200_OK;
}

}  // extern "C"
