// tiles.cu — tile-owner surface voxelization for a PREPARED mesh (voxb200_mesh_*; replaces voxelize.cu:58-238 of the
// reference for meshes that are voxelized more than once, README.md:74, and for the end-to-end path whose upload
// already rewrites the triangles once).
//
// Why: with one global atomicOr per hit row into a table that was cleared by an earlier kernel (surface.cu), a voxelization of
// config 4 moves the 1 GiB table through DRAM three times — zero-fill, line fetch for the atomics, write-back — and the
// zero-fill runs serially in front of the arithmetic.  Here every table line reaches DRAM exactly once:
//
//   prepare (once per mesh and grid, upload path)
//     tile_count_kernel    per triangle: exact grid bbox (the kernel's own arithmetic), class, and a histogram over the
//                          tiles (1024 x 16 x 16 voxels: whole 128-byte lines) its bbox overlaps.  Triangles whose bbox exceeds
//                          4x4x4 voxels ("big") and the triangles of over-full tiles go to a side soup for the row-solver path.
//     tile_plan_kernel     one CTA: offsets, the list of empty tiles, the non-empty tiles heaviest first
//                          (longest-processing-time order for the block scheduler) and their batch prefix.
//     tile_scatter_kernel  writes one 64-byte RECORD per (small triangle, overlapped tile) into the tile's run (1.17 records
//                          per triangle on config 4): the shifted vertices, the unit normal and the plane offsets d1, d2
//                          (the IEEE sqrt / divide part of the setup, computed once per mesh instead of once per call),
//                          and the grid bbox clipped to the tile.
//   voxelize (every call)
//     surface_tile_kernel  one block per non-empty tile: it CLEARS its tile in global memory (32 KB of whole lines, which stay
//                          L2-resident: 888 blocks x 32 KB in flight), fences, and then runs the branch-free <=3x3x3 / <=4x4x4
//                          evaluation of surf_micro.cuh over its records, candidates clipped to the tile (bit-exact by
//                          construction, like the multi-GPU region clipping), ORing hits into the tile with red.or — every
//                          atomic is an L2 hit on a line no other block touches during the kernel, so the table costs one
//                          DRAM write per line and no read.  After every 32-record batch the warp also clears a proportional
//                          share of the EMPTY tiles, so that part of the zero-fill rides inside the arithmetic too.
//     launch_surface(ACCUMULATE) over the side soup: big triangles through the exact row solver, afterwards.
//
// Measured and rejected on the way (B200, config 4, profiles/README.md): tiles assembled in shared memory at one byte per voxel
// (plain idempotent byte stores, no atomics) and packed to bits at the end — 0.82 ms: the tiles a block can hold (256x16x16)
// are written as 32-byte sectors, a quarter of a line each; tiles of 256x32x32 voxels with red.or — 1.47 ms: L2 atomics need
// the whole 128-byte line, so quarter-line tiles fetch the rest from DRAM; 1024x32x32 tiles — 0.64-0.96 ms: 57-114 MB in
// flight do not stay in L2.
//
// The table bits do not depend on any of this: every voxel that is set passed the reference's exact per-voxel
// expression sequence (vox_exact.cuh), evaluated by the same code as the per-triangle kernel.
#include "vox_internal.h"
#include "surf_micro.cuh"

namespace voxb {

constexpr int kPrepBlock = 256;
constexpr unsigned int kKindMicro = 1u << 30, kKindBig = 2u << 30;
constexpr int kLptBins = 1024;

// ------------------------------------------------------------------------------------------------
// prepare
// ------------------------------------------------------------------------------------------------
// Returns false when a vertex index of the face lies outside [0, n_verts) (then clamped: nothing is read out of bounds; the count
// kernel reports the face, voxb200_mesh_create_indexed fails).
template <bool INDEXED>
__device__ __forceinline__ bool load_src_tri(const float* __restrict__ soup, const float* __restrict__ verts, const int* __restrict__ faces,
                                             unsigned long long i, unsigned int n_verts, Tri& t) {
	bool ok = true;
	if (INDEXED) {
		int a = __ldg(faces + 3 * i), b = __ldg(faces + 3 * i + 1), c = __ldg(faces + 3 * i + 2);
		if (n_verts && ((unsigned int)a >= n_verts || (unsigned int)b >= n_verts || (unsigned int)c >= n_verts)) {
			ok = false;
			a = min(max(a, 0), (int)n_verts - 1); b = min(max(b, 0), (int)n_verts - 1); c = min(max(c, 0), (int)n_verts - 1);
		}
		const float* pa = verts + 3 * (size_t)a;
		const float* pb = verts + 3 * (size_t)b;
		const float* pc = verts + 3 * (size_t)c;
		t.v0x = __ldg(pa); t.v0y = __ldg(pa + 1); t.v0z = __ldg(pa + 2);
		t.v1x = __ldg(pb); t.v1y = __ldg(pb + 1); t.v1z = __ldg(pb + 2);
		t.v2x = __ldg(pc); t.v2y = __ldg(pc + 1); t.v2z = __ldg(pc + 2);
	} else {
		load_tri_aos(soup, i, t);
	}
	return ok;
}

__device__ __forceinline__ unsigned int tile_id(const TileGeom& tg, int tx, int ty, int tzl) {
	return ((unsigned int)tzl * (unsigned int)tg.nty + (unsigned int)ty) * (unsigned int)tg.ntx + (unsigned int)tx;
}

// The region-clipped grid bbox of a (shifted) triangle; false when it misses the region.
__device__ __forceinline__ bool region_bbox(const Tri& t, const GridParams& g, SurfSetup& s) {
	surf_bbox(t, g, s);
	s.x0 = max(s.x0, g.rx0); s.x1 = min(s.x1, g.rx1 - 1);
	s.y0 = max(s.y0, g.ry0); s.y1 = min(s.y1, g.ry1 - 1);
	s.z0 = max(s.z0, g.rz0); s.z1 = min(s.z1, g.rz1 - 1);
	return s.x0 <= s.x1 && s.y0 <= s.y1 && s.z0 <= s.z1;
}

// key: bits 0..19 tile of the bbox's min corner, bit 20/21/22 the bbox continues into the next tile along x/y/z, bit 23 the bbox is
// 4 voxels long on some axis, bits 30..31 kind
template <bool INDEXED>
__global__ void __launch_bounds__(kPrepBlock) tile_count_kernel(const GridParams g, const TileGeom tg, const float* __restrict__ soup,
                                                                const float* __restrict__ verts, const int* __restrict__ faces,
                                                                unsigned int* __restrict__ keys, unsigned int* __restrict__ cnt,
                                                                unsigned long long* __restrict__ totals) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kPrepBlock + threadIdx.x;
	unsigned int key = 0u;
	bool bad = false;
	if (i < g.n_tris) {
		Tri t;
		bad = !load_src_tri<INDEXED>(soup, verts, faces, i, tg.n_verts, t);
		shift_tri(t, g);
		SurfSetup s;
		if (region_bbox(t, g, s)) {
			const bool micro = s.x1 - s.x0 <= 3 && s.y1 - s.y0 <= 3 && s.z1 - s.z0 <= 3;
			if (micro) {
				const int tx0 = s.x0 >> tg.tx_shift, ty0 = s.y0 / kTileY, tz0 = s.z0 / kTileZ - tg.tz0;
				const unsigned int sx = ((s.x1 >> tg.tx_shift) != tx0), sy = (s.y1 / kTileY != ty0), sz = (s.z1 / kTileZ - tg.tz0 != tz0);
				const unsigned int wide = s.x1 - s.x0 == 3 || s.y1 - s.y0 == 3 || s.z1 - s.z0 == 3;      // needs the 4x4x4 evaluation
				key = kKindMicro | tile_id(tg, tx0, ty0, tz0) | (sx << 20) | (sy << 21) | (sz << 22) | (wide << 23);
			} else {
				key = kKindBig;
			}
		}
		keys[i] = key;
	}
	// histogram: the min-corner tile warp-aggregated (neighbouring triangles of a mesh mostly share it), the continuation tiles directly
	const bool is_micro = (key >> 30) == 1u;
	const unsigned int act = __ballot_sync(0xffffffffu, is_micro);
	if (is_micro) {
		const unsigned int t0 = key & 0xfffffu;
		const unsigned int peers = __match_any_sync(act, t0);
		if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(cnt + t0, (unsigned int)__popc(peers));
		const unsigned int sx = (key >> 20) & 1u, sy = (key >> 21) & 1u, sz = (key >> 22) & 1u;
		if (sx | sy | sz) {
			for (unsigned int m = 1; m < 8; m++) {
				if ((m & 1u) > sx || ((m >> 1) & 1u) > sy || ((m >> 2) & 1u) > sz) continue;
				atomicAdd(cnt + t0 + (m & 1u) + (unsigned int)tg.ntx * (((m >> 1) & 1u) + (unsigned int)tg.nty * ((m >> 2) & 1u)), 1u);
			}
		}
	}
	const unsigned int bigs = __ballot_sync(0xffffffffu, (key >> 30) == 2u);
	if (bigs && (threadIdx.x & 31) == 0) atomicAdd(totals + kPlanBigDirect, (unsigned long long)__popc(bigs));
	if (INDEXED) {
		const unsigned int bads = __ballot_sync(0xffffffffu, bad);
		if (bads && (threadIdx.x & 31) == 0) atomicAdd(totals + kPlanBadFaces, (unsigned long long)__popc(bads));
	}
	const unsigned int wides = __ballot_sync(0xffffffffu, (key >> 23) & 1u);
	if (wides && (threadIdx.x & 31) == 0) atomicAdd(totals + kPlanWide, (unsigned long long)__popc(wides));
}

// Exclusive scan of one value per thread over a block of 1024 threads (shuffles inside the warps, then the 32 warp totals);
// returns the exclusive prefix, `total` = the block's sum.  Contains two barriers.
template <typename T>
__device__ __forceinline__ T block_scan_1024(T v, T* warp_totals, T& total) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	T inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const T up = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += up;
	}
	if (lane == 31) warp_totals[wid] = inc;
	__syncthreads();
	const T wt = warp_totals[lane];
	T winc = wt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const T up = __shfl_up_sync(0xffffffffu, winc, d);
		if (lane >= d) winc += up;
	}
	const T warp_base = __shfl_sync(0xffffffffu, winc - wt, wid);
	total = __shfl_sync(0xffffffffu, winc, 31);
	__syncthreads();
	return warp_base + inc - v;
}

// One CTA of 1024 threads plans the whole grid of tiles (n_tiles <= 2^20): see the header comment.  Every thread owns a
// contiguous chunk of the tiles (so offsets and the empty-tile list come out in table order); prefixes across the chunks are block
// scans.  The work records are staged in place — pass 2 leaves {tile, records, first record} in the slot the tile takes in the
// heaviest-first order, the last pass completes them — so no pass chases a pointer (the first version, with thread-0 loops and
// cnt[order[w]] gathers, took 0.167 ms for the 32,768 tiles of config 4, as long as the scatter of 12 M records it plans).
__global__ void __launch_bounds__(1024) tile_plan_kernel(const TileGeom tg, const unsigned int cap, const unsigned int* __restrict__ cnt,
                                                         unsigned int* __restrict__ off, unsigned int* __restrict__ order,
                                                         uint4* __restrict__ work, unsigned int* __restrict__ empty,
                                                         unsigned long long* __restrict__ totals, unsigned int* __restrict__ scratch) {
	__shared__ unsigned long long s_wt64[32];
	__shared__ unsigned int s_wt32[32];
	__shared__ unsigned int s_bin[kLptBins], s_cursor[kLptBins];
	__shared__ unsigned long long s_heavy;
	const unsigned int n = tg.n_tiles;
	const unsigned int per = (n + 1023u) / 1024u;
	const unsigned int a = min(n, threadIdx.x * per), b = min(n, a + per);
	s_bin[threadIdx.x] = 0u; s_cursor[threadIdx.x] = 0u;       // kLptBins == blockDim.x
	if (threadIdx.x == 0) s_heavy = 0ull;
	__syncthreads();
	unsigned long long inst = 0ull, heavy = 0ull;
	unsigned int n_empty = 0u, n_work = 0u;
#pragma unroll 8
	for (unsigned int t = a; t < b; t++) {
		const unsigned int raw = cnt[t];
		const unsigned int c = raw > cap ? 0u : raw;
		if (raw > cap) heavy += raw;
		inst += c;
		if (c == 0u) n_empty++;
		else { n_work++; atomicAdd(&s_bin[min(c >> 3, (unsigned int)kLptBins - 1u)], 1u); }
	}
	if (heavy) atomicAdd(&s_heavy, heavy);
	unsigned long long total_inst;
	unsigned int total_empty, total_work;
	unsigned long long ri = block_scan_1024<unsigned long long>(inst, s_wt64, total_inst);
	unsigned int re = block_scan_1024<unsigned int>(n_empty, s_wt32, total_empty);
	block_scan_1024<unsigned int>(n_work, s_wt32, total_work);
	// heaviest bin first: bin k starts behind all heavier bins
	unsigned int bins_total;
	const unsigned int rev = (unsigned int)kLptBins - 1u - threadIdx.x;
	const unsigned int bin_start = block_scan_1024<unsigned int>(s_bin[rev], s_wt32, bins_total);
	s_bin[rev] = bin_start;
	if (threadIdx.x == 0) {
		totals[kPlanInstances] = total_inst; totals[kPlanEmpty] = total_empty; totals[kPlanWork] = total_work; totals[kPlanHeavyInstances] = s_heavy;
		off[n] = (unsigned int)min(total_inst, 0xffffffffull);
	}
	__syncthreads();
	const unsigned long long G = (unsigned long long)tg.G;
#pragma unroll 4
	for (unsigned int t = a; t < b; t++) {
		const unsigned int raw = cnt[t];
		const unsigned int c = raw > cap ? 0u : raw;
		off[t] = (unsigned int)ri;
		if (c == 0u) {
			// the table word (relative to the region) of the tile's first voxel: the region holds at most 2^32 words (mesh_tileable)
			const unsigned int tx = t % (unsigned int)tg.ntx, r = t / (unsigned int)tg.ntx;
			const unsigned int ty = r % (unsigned int)tg.nty, tzl = r / (unsigned int)tg.nty;
			empty[re++] = (unsigned int)((((unsigned long long)tzl * kTileZ * G + (unsigned long long)ty * kTileY) * G + ((unsigned long long)tx << tg.tx_shift)) >> 5);
		} else {
			const unsigned int bin = min(c >> 3, (unsigned int)kLptBins - 1u);
			const unsigned int slot = s_bin[bin] + atomicAdd(&s_cursor[bin], 1u);
			order[slot] = t;
			work[slot] = make_uint4(t, c, (unsigned int)ri, 0u);           // completed below
		}
		ri += c;
	}
	__threadfence_block();
	__syncthreads();
	// batch prefix in work order, then the finished records
	const unsigned int nw = total_work;
	const unsigned int perw = (nw + 1023u) / 1024u;
	const unsigned int wa = min(nw, threadIdx.x * perw), wb = min(nw, wa + perw);
	unsigned int nb = 0u;
#pragma unroll 8
	for (unsigned int w = wa; w < wb; w++) nb += (work[w].y + 31u) >> 5;
	unsigned int total_batches;
	unsigned int run = block_scan_1024<unsigned int>(nb, s_wt32, total_batches);
	if (threadIdx.x == 0) totals[kPlanBatches] = total_batches;
#pragma unroll 4
	for (unsigned int w = wa; w < wb; w++) {
		// everything a tile block needs, in one 16-byte load: {table word of the tile's first voxel, records, first record, batches before it}
		const uint4 st = work[w];
		const unsigned int t = st.x, c = st.y;
		const unsigned int tx = t % (unsigned int)tg.ntx, r = t / (unsigned int)tg.ntx;
		const unsigned int ty = r % (unsigned int)tg.nty, tzl = r / (unsigned int)tg.nty;
		const unsigned int word = (unsigned int)((((unsigned long long)tzl * kTileZ * G + (unsigned long long)ty * kTileY) * G + ((unsigned long long)tx << tg.tx_shift)) >> 5);
		work[w] = make_uint4(word, c, st.z, run);
		run += (c + 31u) >> 5;
	}
	(void)scratch;
}

// The 64-byte record of one (triangle, tile) pair:
//   words 0..8   the vertices, already shifted by -bbox.min (cpu_voxelizer.cpp:40-45)
//   words 9..11  the unit normal (:72), words 12, 13 the plane offsets d1, d2 (:83-87)
//   word 14      x0 | y0 << 16, word 15  z0 | ex << 16 | ey << 18 | ez << 20: the grid bbox clipped to region and tile (extents - 1)
template <bool INDEXED>
__global__ void __launch_bounds__(kPrepBlock) tile_scatter_kernel(const GridParams g, const TileGeom tg, const unsigned int cap,
                                                                  const float* __restrict__ soup, const float* __restrict__ verts,
                                                                  const int* __restrict__ faces, const unsigned int* __restrict__ keys,
                                                                  const unsigned int* __restrict__ cnt, const unsigned int* __restrict__ off,
                                                                  unsigned int* __restrict__ fill, uint4* __restrict__ records,
                                                                  float* __restrict__ side, unsigned long long* __restrict__ totals) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kPrepBlock + threadIdx.x;
	const unsigned int key = i < g.n_tris ? keys[i] : 0u;
	const unsigned int kind = key >> 30;
	Tri t;
	if (kind) load_src_tri<INDEXED>(soup, verts, faces, i, tg.n_verts, t);
	bool to_side = kind == 2u;
	const unsigned int act = __ballot_sync(0xffffffffu, kind == 1u);
	if (kind == 1u) {
		Tri ts = t;
		shift_tri(ts, g);
		SurfSetup s;
		region_bbox(ts, g, s);
		surf_setup_tests<true>(ts, g, s);            // the normal and d1, d2 are what is kept
		const unsigned int t0 = key & 0xfffffu;
		const unsigned int sx = (key >> 20) & 1u, sy = (key >> 21) & 1u, sz = (key >> 22) & 1u;
		const int tx0 = s.x0 >> tg.tx_shift, ty0 = s.y0 / kTileY, tz0 = s.z0 / kTileZ;
		const int lane = threadIdx.x & 31;
		for (unsigned int m = 0; m < 8; m++) {
			const unsigned int mx = m & 1u, my = (m >> 1) & 1u, mz = (m >> 2) & 1u;
			if (mx > sx || my > sy || mz > sz) continue;
			const unsigned int tile = t0 + mx + (unsigned int)tg.ntx * (my + (unsigned int)tg.nty * mz);
			const bool heavy = __ldg(cnt + tile) > cap;                        // over-full tile: its triangles take the side path
			unsigned int pos = 0u;
			if (m == 0u) {
				// the min-corner tile: one cursor bump per group of lanes that share it (every micro lane gets here)
				const unsigned int peers = __match_any_sync(act, t0);
				const int leader = __ffs(peers) - 1;
				unsigned int base = 0u;
				if (lane == leader && !heavy) base = atomicAdd(fill + tile, (unsigned int)__popc(peers));
				base = __shfl_sync(peers, base, leader);
				pos = base + (unsigned int)__popc(peers & ((1u << lane) - 1u));
			} else if (!heavy) {
				pos = atomicAdd(fill + tile, 1u);
			}
			if (heavy) { to_side = true; continue; }
			// the bbox inside this tile
			const int bx0 = (tx0 + (int)mx) << tg.tx_shift, by0 = (ty0 + (int)my) * kTileY, bz0 = (tz0 + (int)mz) * kTileZ;
			const int x0 = max(s.x0, bx0), x1 = min(s.x1, bx0 + (1 << tg.tx_shift) - 1);
			const int y0 = max(s.y0, by0), y1 = min(s.y1, by0 + kTileY - 1);
			const int z0 = max(s.z0, bz0), z1 = min(s.z1, bz0 + kTileZ - 1);
			uint4* o = records + 4ull * ((unsigned long long)__ldg(off + tile) + pos);
			o[0] = make_uint4(__float_as_uint(ts.v0x), __float_as_uint(ts.v0y), __float_as_uint(ts.v0z), __float_as_uint(ts.v1x));
			o[1] = make_uint4(__float_as_uint(ts.v1y), __float_as_uint(ts.v1z), __float_as_uint(ts.v2x), __float_as_uint(ts.v2y));
			o[2] = make_uint4(__float_as_uint(ts.v2z), __float_as_uint(s.nx), __float_as_uint(s.ny), __float_as_uint(s.nz));
			o[3] = make_uint4(__float_as_uint(s.d1), __float_as_uint(s.d2), (unsigned int)x0 | ((unsigned int)y0 << 16),
			                  (unsigned int)z0 | ((unsigned int)(x1 - x0) << 16) | ((unsigned int)(y1 - y0) << 18) | ((unsigned int)(z1 - z0) << 20));
		}
	}
	const unsigned int m = __ballot_sync(0xffffffffu, to_side);
	if (m == 0u) return;
	const int lane = threadIdx.x & 31;
	unsigned long long base = 0ull;
	if (lane == 0) base = atomicAdd(totals + kPlanSideFill, (unsigned long long)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (to_side) {
		float* o = side + 9ull * (base + __popc(m & ((1u << lane) - 1u)));
		o[0] = t.v0x; o[1] = t.v0y; o[2] = t.v0z; o[3] = t.v1x; o[4] = t.v1y; o[5] = t.v1z; o[6] = t.v2x; o[7] = t.v2y; o[8] = t.v2z;
	}
}

cudaError_t launch_tile_count(const GridParams& g, const TileGeom& tg, const float* d_soup, const float* d_verts, const int* d_faces,
                              unsigned int* d_keys, unsigned int* d_cnt, unsigned long long* d_totals, cudaStream_t st) {
	cudaError_t e = cudaMemsetAsync(d_cnt, 0, (size_t)tg.n_tiles * sizeof(unsigned int), st);
	if (e == cudaSuccess) e = cudaMemsetAsync(d_totals, 0, kPlanTotals * sizeof(unsigned long long), st);
	if (e != cudaSuccess || g.n_tris == 0) return e;
	const unsigned int blocks = (unsigned int)((g.n_tris + kPrepBlock - 1) / kPrepBlock);
	if (d_faces) tile_count_kernel<true><<<blocks, kPrepBlock, 0, st>>>(g, tg, nullptr, d_verts, d_faces, d_keys, d_cnt, d_totals);
	else tile_count_kernel<false><<<blocks, kPrepBlock, 0, st>>>(g, tg, d_soup, nullptr, nullptr, d_keys, d_cnt, d_totals);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_tile_plan(const TileGeom& tg, unsigned int cap, const unsigned int* d_cnt, unsigned int* d_off, unsigned int* d_order,
                             void* d_work, unsigned int* d_empty, unsigned long long* d_totals, unsigned int* d_scratch, cudaStream_t st) {
	tile_plan_kernel<<<1, 1024, 0, st>>>(tg, cap, d_cnt, d_off, d_order, reinterpret_cast<uint4*>(d_work), d_empty, d_totals, d_scratch);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_tile_scatter(const GridParams& g, const TileGeom& tg, unsigned int cap, const float* d_soup, const float* d_verts,
                                const int* d_faces, const unsigned int* d_keys, const unsigned int* d_cnt, const unsigned int* d_off,
                                unsigned int* d_fill, void* d_records, float* d_side, unsigned long long* d_totals, cudaStream_t st) {
	cudaError_t e = cudaMemsetAsync(d_fill, 0, (size_t)tg.n_tiles * sizeof(unsigned int), st);
	if (e != cudaSuccess || g.n_tris == 0) return e;
	const unsigned int blocks = (unsigned int)((g.n_tris + kPrepBlock - 1) / kPrepBlock);
	uint4* rec = reinterpret_cast<uint4*>(d_records);
	if (d_faces) tile_scatter_kernel<true><<<blocks, kPrepBlock, 0, st>>>(g, tg, cap, nullptr, d_verts, d_faces, d_keys, d_cnt, d_off, d_fill, rec, d_side, d_totals);
	else tile_scatter_kernel<false><<<blocks, kPrepBlock, 0, st>>>(g, tg, cap, d_soup, nullptr, nullptr, d_keys, d_cnt, d_off, d_fill, rec, d_side, d_totals);
	g_launch_count++;
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// voxelize
// ------------------------------------------------------------------------------------------------
// Chunk c (512 bytes) of the empty space — the empty tiles back to back; p.empty holds the table word of each one's first voxel.
// A chunk is 4096 >> tx_shift consecutive row segments (y fastest), Tx / 128 lanes of 16 bytes per segment.
__device__ __forceinline__ void zero_chunk_at(const GridParams& g, const TilePlan& p, unsigned int* __restrict__ table, unsigned int base, unsigned int c, int lane) {
	const unsigned int sub = c & ((1u << p.geom.chunk_shift) - 1u);
	const int lanes_shift = p.geom.tx_shift - 7;
	const unsigned int row = (sub << (5 - lanes_shift)) + ((unsigned int)lane >> lanes_shift);
	const unsigned int Gw = (unsigned int)g.G >> 5;
	const unsigned int word = base + ((row / (unsigned int)kTileY) * (unsigned int)g.G + (row % (unsigned int)kTileY)) * Gw + 4u * ((unsigned int)lane & ((1u << lanes_shift) - 1u));
	*reinterpret_cast<uint4*>(table + word) = make_uint4(0u, 0u, 0u, 0u);
}
#ifndef VOXB_RED_BLOCK
#define VOXB_RED_BLOCK 128
#endif
#ifndef VOXB_RED_MINB
#define VOXB_RED_MINB 6
#endif
constexpr int kRedBlock = VOXB_RED_BLOCK;
constexpr int kRecWords = 16;                          // 64-byte records
constexpr int kBatchVec = 32 * kRecWords / 4;          // one batch = 2 KB = 128 x 16 bytes

// 16-byte piece k of record r of a staged batch sits at piece index 4r + (k ^ ((r >> 1) & 3)): the four LDS.128 of a lane's record
// are then conflict-free (eight consecutive lanes cover all eight 16-byte bank groups).
__device__ __forceinline__ void fetch_batch(const uint4* __restrict__ src, uint4* dst, int lane) {
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const int q = j * 32 + lane, r = q >> 2, k = (q & 3) ^ ((r >> 1) & 3);
		cp_async16(dst + q, src + 4 * r + k);
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}

// WIDE = some record of the plan needs the 64-candidate (<=4x4x4) evaluation: that variant gives up occupancy for registers.
template <bool ACC, bool WIDE>
__global__ void __launch_bounds__(kRedBlock, WIDE ? 4 : VOXB_RED_MINB) surface_tile_kernel(const GridParams g, const TilePlan p, unsigned int* __restrict__ table) {
	__shared__ __align__(16) uint4 stage[(kRedBlock / 32) * 2 * kBatchVec];           // two batches per warp
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	grid_launch_dependents();                // the side path's kernels (big triangles) may be scheduled during this grid's last wave
	if (blockIdx.x < p.n_zero_blocks) {
		// the part of the empty space that is not cleared by the tile blocks below
		const unsigned int c0 = p.zero_rest_first + blockIdx.x * (unsigned int)kZeroBlockChunks;
		const unsigned int c1 = min(p.zero_chunks, c0 + (unsigned int)kZeroBlockChunks);
		unsigned int cur = ~0u, base = 0u;             // the tile address is loaded once per tile, so the stores do not wait on it
#pragma unroll 4
		for (unsigned int c = c0 + warp; c < c1; c += kRedBlock / 32) {
			const unsigned int tile = c >> p.geom.chunk_shift;
			if (tile != cur) { base = __ldg(p.empty + tile); cur = tile; }
			zero_chunk_at(g, p, table, base, c, lane);
		}
		return;
	}
	const uint4 wk = __ldg(p.work + (blockIdx.x - p.n_zero_blocks));       // {tile's first table word, records, first record, batches before}
	const unsigned int cnt = wk.y, nb = (wk.y + 31u) >> 5, bp = wk.w;
	const uint4* recs = reinterpret_cast<const uint4*>(p.soup) + 4ull * wk.z;
	uint4* my_stage = stage + warp * 2 * kBatchVec;
	if ((unsigned int)warp < nb) fetch_batch(recs + (size_t)warp * kBatchVec, my_stage, lane);        // in flight during the clear
	if (!ACC) {
		// clear the tile: a thread's 16-byte pieces are a whole number of z-layers apart, so one address and a constant stride
		const int lanes_shift = p.geom.tx_shift - 7;                                       // lanes per row segment = Tx / 128
		const unsigned int row = threadIdx.x >> lanes_shift;
		const unsigned int Gw = (unsigned int)g.G >> 5;
		uint4* dst = reinterpret_cast<uint4*>(table + (wk.x + ((row / (unsigned int)kTileY) * (unsigned int)g.G + (row % (unsigned int)kTileY)) * Gw)) +
		             (threadIdx.x & ((1u << lanes_shift) - 1u));
		const int layers_per_step = (kRedBlock >> lanes_shift) / kTileY;                  // the block's threads cover this many z-layers per step
		const size_t stride = (size_t)layers_per_step * ((size_t)g.G * (size_t)g.G / 128u);      // in 16-byte units
		for (int k = 0; k < kTileZ / layers_per_step; k++) dst[(size_t)k * stride] = make_uint4(0u, 0u, 0u, 0u);
	}
	// The clears must be ordered before every red.or of the block into the tile.  Both come from threads of THIS block and no other
	// block touches the tile during the kernel, so a block barrier is all the ordering needed (accesses made before bar.sync are
	// visible to every thread of the block after it, and an atomic acts on the visible value; a device-scope fence would only add a
	// store drain: it cost 10 % of the kernel when it was here).  The barrier sits in front of the first SCATTER, not the first
	// batch, so nobody waits for the clears before there is something to write.
	bool unfenced = !ACC;
	int buf = 0;
#pragma unroll 1
	for (unsigned int b = warp; b < nb; b += kRedBlock / 32) {
		// this batch's share of the empty space: the tile address now, the stores after the arithmetic
		unsigned int c0 = 0u, c1 = 0u, zbase = 0u;
		if (!ACC) {
			c0 = min(p.zero_chunks, (bp + b) * p.zero_quota); c1 = min(p.zero_chunks, c0 + p.zero_quota);
			if (c0 < c1) zbase = __ldg(p.empty + (c0 >> p.geom.chunk_shift));
		}
		// the next batch's records while this batch's arrive
		const unsigned int bn = b + kRedBlock / 32;
		if (bn < nb) fetch_batch(recs + (size_t)bn * kBatchVec, my_stage + (buf ^ 1) * kBatchVec, lane);
		if (bn < nb) asm volatile("cp.async.wait_group 1;" ::: "memory");
		else asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		const uint4* rec = my_stage + buf * kBatchVec + 4 * lane;
		const int sw = (lane >> 1) & 3;
		const uint4 r0 = rec[0 ^ sw], r1 = rec[1 ^ sw], r2 = rec[2 ^ sw], r3 = rec[3 ^ sw];
		__syncwarp();                  // the stage buffer is refilled two iterations from now: every lane has read its record by then
		buf ^= 1;
		const bool live = 32u * b + (unsigned int)lane < cnt;
		Tri tr;
		tr.v0x = __uint_as_float(r0.x); tr.v0y = __uint_as_float(r0.y); tr.v0z = __uint_as_float(r0.z); tr.v1x = __uint_as_float(r0.w);
		tr.v1y = __uint_as_float(r1.x); tr.v1z = __uint_as_float(r1.y); tr.v2x = __uint_as_float(r1.z); tr.v2y = __uint_as_float(r1.w);
		tr.v2z = __uint_as_float(r2.x);
		SurfSetup s;
		s.nx = __uint_as_float(r2.y); s.ny = __uint_as_float(r2.z); s.nz = __uint_as_float(r2.w);
		s.d1 = __uint_as_float(r3.x); s.d2 = __uint_as_float(r3.y);
		const int ex = (int)((r3.w >> 16) & 3u), ey = (int)((r3.w >> 18) & 3u), ez = (int)((r3.w >> 20) & 3u);
		s.x0 = (int)(r3.z & 0xffffu); s.y0 = (int)(r3.z >> 16); s.z0 = (int)(r3.w & 0xffffu);
		s.x1 = s.x0 + ex; s.y1 = s.y0 + ey; s.z1 = s.z0 + ez;
		if (live) surf_setup_tests<false>(tr, g, s);
		const bool wide = WIDE && __any_sync(0xffffffffu, live && (ex == 3 || ey == 3 || ez == 3));
		unsigned long long hit4 = 0ull;
		unsigned int hit = 0u;
		if (wide) hit4 = live ? surf_micro4(s, g) : 0ull;
		else hit = live ? surf_micro3(s, g) : 0u;
		if (unfenced) {
			asm volatile("bar.sync 1, %0;" ::"n"(kRedBlock) : "memory");
			unfenced = false;
		}
		if (wide) {
			if (hit4) scatter_hits4<false>(hit4, s.x0, s.y0, s.z0, g, table);
		} else {
			if (!live) { s.x0 = g.rx0; s.y0 = g.ry0; s.z0 = g.rz0; }
			scatter_hits3<false>(hit, s.x0, s.y0, s.z0, g, table);      // converged: every write is predicated on its own bits
		}
		// (the empty tiles are nobody's: no ordering needed)
		for (unsigned int c = c0; c < c1; c++) {
			const unsigned int base = (c >> p.geom.chunk_shift) == (c0 >> p.geom.chunk_shift) ? zbase : __ldg(p.empty + (c >> p.geom.chunk_shift));
			zero_chunk_at(g, p, table, base, c, lane);
		}
	}
	if (unfenced) asm volatile("bar.sync 1, %0;" ::"n"(kRedBlock) : "memory");          // a warp without batches still owes its arrival
}

template <bool ACC, bool WIDE>
static cudaError_t run_tiles(const GridParams& g, const TilePlan& p, unsigned int* d_table, unsigned int blocks, cudaStream_t st) {
	surface_tile_kernel<ACC, WIDE><<<blocks, kRedBlock, 0, st>>>(g, p, d_table);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_surface_tiles(const GridParams& g, const TilePlan& p, unsigned int* d_table, bool accumulate, cudaStream_t st) {
	const unsigned int blocks = p.n_work + (accumulate ? 0u : p.n_zero_blocks);
	if (blocks == 0) return cudaSuccess;
	if (accumulate) {
		TilePlan q = p;
		q.n_zero_blocks = 0;
		return p.wide ? run_tiles<true, true>(g, q, d_table, blocks, st) : run_tiles<true, false>(g, q, d_table, blocks, st);
	}
	return p.wide ? run_tiles<false, true>(g, p, d_table, blocks, st) : run_tiles<false, false>(g, p, d_table, blocks, st);
}

}  // namespace voxb
