// surface.cu — surface voxelization for sm_100a (replaces voxelize.cu:58-238 of the reference).
//
// Schedule (one stream, no host synchronisation anywhere):
//   zero_kernel          the region's table bytes, 16-byte stores                       (skipped with ACCUMULATE)
//   surface_tri_kernel   one warp per tile of 32 consecutive triangles.  The tile's 1152 bytes come in with
//                        16-byte cp.async copies into a warp-private shared slab, each lane takes one triangle:
//                        exact setup, grid bbox, then
//                          - bbox <= 3x3x3 ("micro", the regime of meshes tessellated near the voxel size):
//                            all 27 candidates in branch-free straight-line code -> 27-bit hit mask, written
//                            one (y,z) row (3 x-adjacent bits) per atomicOr; bbox <= 4x4x4: the same with 64
//                            candidates, for the whole warp as soon as one of its triangles needs it;
//                          - anything bigger: queued for the cooperative kernel ({slot, work items} reserved
//                            with ONE packed 64-bit atomic per warp).
//   surface_coop_kernel  persistent grid over the queued (y,z) rows, 64 consecutive rows per warp and iteration;
//                        8 lanes per item, ONE LANE PER ROW.  A row failing the x-independent YZ tests is dropped;
//                        otherwise the lane SOLVES the row instead of sweeping it: every remaining test is a
//                        monotone function of x (rounding, int->float and multiplying/adding a constant all
//                        preserve order), so the accepted voxels of a row form one interval whose two ends are
//                        found by evaluating the reference's exact expressions at a real-arithmetic estimate
//                        (bisection when the estimate is off) — O(1)..O(log width) evaluations instead of width —
//                        and written as whole-word masks.  Setups come from the per-triangle kernel (160 B each).
//
// Measured and rejected on B200 (10M-triangle mesh @2048^3, profiles/README.md): a persistent per-triangle grid
// (static stride or ticket counter: +50 % time, DRAM re-reads double); parking results while the zero-fill
// drains and writing them from a second launch (the atomics, no longer hidden under arithmetic, cost what the
// overlap saved); staging table words in a shared-memory hash per block (+65 %).
//
// Every voxel that is set passed the reference's exact per-voxel expression sequence (vox_exact.cuh).
#include "vox_internal.h"
#include "surf_micro.cuh"

namespace voxb {

constexpr int kBlock = 256;        // cooperative kernel
constexpr int kTriBlock = 128;     // per-triangle kernel: 4 warps, each with its own staging slab
constexpr int kSweepWidth = 8;     // cooperative kernel: rows narrower than this are swept voxel by voxel instead of solved
constexpr int kRowsPerWarp = 64;   // consecutive queued (y,z) rows one warp of the cooperative kernel takes per iteration (= one directory bucket)
#ifndef VOXB_TRI_MINBLOCKS
#define VOXB_TRI_MINBLOCKS 6       // per-triangle kernel: 6 blocks of 128 per SM (<= 80 registers)
#endif
#ifndef VOXB_DEBUG_RED_MODE
#define VOXB_DEBUG_RED_MODE 0
#endif
#ifndef VOXB_DEBUG_WINDOW_WORDS
#define VOXB_DEBUG_WINDOW_WORDS 0x400000u
#endif
#if VOXB_TRI_MINBLOCKS > 0
#define VOXB_TRI_BOUNDS __launch_bounds__(kTriBlock, VOXB_TRI_MINBLOCKS)
#else
#define VOXB_TRI_BOUNDS __launch_bounds__(kTriBlock)
#endif

std::atomic<unsigned long long> g_launch_count{0};

// ------------------------------------------------------------------------------------------------
// Both also reset the library's per-call counters (when given), which saves a separate memset node per call.
__global__ void __launch_bounds__(512) zero_kernel(uint4* __restrict__ p, size_t n16, unsigned long long* __restrict__ counters) {
	grid_launch_dependents();
	if (counters != nullptr && blockIdx.x == 0 && threadIdx.x < kNumCounters) counters[threadIdx.x] = 0ull;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(512) zero_words_kernel(unsigned int* __restrict__ p, size_t n, unsigned long long* __restrict__ counters) {
	grid_launch_dependents();
	if (counters != nullptr && blockIdx.x == 0 && threadIdx.x < kNumCounters) counters[threadIdx.x] = 0ull;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0u;
}

cudaError_t launch_zero(Workspace& ws, unsigned int* p, size_t words, cudaStream_t st, bool reset_counters) {
	unsigned long long* counters = reset_counters ? ws.counters : nullptr;
	if (words == 0) return reset_counters ? cudaMemsetAsync(ws.counters, 0, kNumCounters * sizeof(unsigned long long), st) : cudaSuccess;
	const int sms = ws.sm_count > 0 ? ws.sm_count : 148;
	if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (words & 3u) == 0) {
		const size_t n16 = words / 4;
		size_t blocks = (n16 + 511) / 512;
		if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
		zero_kernel<<<(unsigned)blocks, 512, 0, st>>>(reinterpret_cast<uint4*>(p), n16, counters);
	} else {
		size_t blocks = (words + 511) / 512;
		if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
		zero_words_kernel<<<(unsigned)blocks, 512, 0, st>>>(p, words, counters);
	}
	g_launch_count++;
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool clip_to_region(const GridParams& g, SurfSetup& s) {
	s.x0 = max(s.x0, g.rx0); s.x1 = min(s.x1, g.rx1 - 1);
	s.y0 = max(s.y0, g.ry0); s.y1 = min(s.y1, g.ry1 - 1);
	s.z0 = max(s.z0, g.rz0); s.z1 = min(s.z1, g.rz1 - 1);
	return s.x0 <= s.x1 && s.y0 <= s.y1 && s.z0 <= s.z1;
}

// ------------------------------------------------------------------------------------------------
// Micro path: a triangle whose (region-clipped) grid bbox is at most 3x3x3 voxels — every triangle of
// a mesh tessellated near the voxel size.  All 27 candidates are evaluated in straight-line code with
// no data-dependent branch, so the 32 triangles of a warp stay converged:
//   * each test value is the reference's exact expression (same products, same left-to-right adds);
//     the 2D edge functions are evaluated once per (x,y) / (y,z) / (z,x) cell instead of once per voxel;
//   * a test's outcome is the SIGN BIT of its value, shifted into a reject mask with one funnel shift:
//       edge test  "v < 0"          -> sign(min(v_e0, v_e1, v_e2))   (v is never -0: its last addend d_e never is)
//       plane test "(s+d1)(s+d2) > 0" -> sign(0 - P)                 (0 - P is -0 for no P; NaN is canonical, sign clear)
//   * voxel b = (2-i) + 3j + 9k (x0 at the high bit of each 3-bit row, MSB-first like the table) survives iff it is inside the bbox and none of the four masks rejects it.
// Survivors are written one (y,z) row at a time: the 3 x-bits of a row go out as one or two atomicOr.
// ------------------------------------------------------------------------------------------------
// One warp = one tile of 32 consecutive triangles.
template <bool MORTON, bool SOA4>
__device__ __forceinline__ void tri_tile(const GridParams& g, const float* __restrict__ tris, unsigned int* __restrict__ table,
                                         const QueueView& q, unsigned long long tile, float* my_stage) {
	const int lane = threadIdx.x & 31;
	Tri t;
	const bool valid = load_tile_tri<SOA4>(g, tris, tile, lane, my_stage, t);
	const unsigned long long i = (tile << 5) + lane;

	SurfSetup s;
	bool live = false, big = false, micro = false;
	unsigned int items = 0u;
	if (valid) {
		shift_tri(t, g);
		surf_bbox(t, g, s);
		live = clip_to_region(g, s);          // triangles of other slabs (multi-GPU regions) stop here
	}
	if (!__any_sync(0xffffffffu, live)) return;
	bool fits3 = false;
	if (live) {
		surf_setup_tests(t, g, s);
		const int dx = s.x1 - s.x0, dy = s.y1 - s.y0, dz = s.z1 - s.z0;
		fits3 = dx <= 2 && dy <= 2 && dz <= 2;
		micro = dx <= 3 && dy <= 3 && dz <= 3;
		const unsigned long long rows = (unsigned long long)(dy + 1) * (unsigned long long)(dz + 1);
		big = !micro;
		items = (unsigned int)rows;          // work units of the cooperative kernel = (y,z) rows (G <= 65536: fits 32 bits)
	}
	const unsigned int slot = enqueue_warp(live && big, items, (unsigned int)i, q);
	if (slot < q.setup_cap) store_setup(q.setups + (size_t)slot * kSetupVec, s);
	const bool mine = live && !big;
	const bool wide = __any_sync(0xffffffffu, mine && !fits3);      // one code path per warp: 64 candidates as soon as one triangle needs them
	unsigned long long hit4 = 0ull;
	unsigned int hit = 0u;
	if (wide) { if (mine) hit4 = surf_micro4(s, g); }
	else if (mine) hit = surf_micro3(s, g);
	if (wide) {
		if (hit4) scatter_hits4<MORTON>(hit4, s.x0, s.y0, s.z0, g, table);
		return;
	}
	if (!MORTON && (g.G & 31) == 0 && g.w32) {
		if (!mine) { s.x0 = 0; s.y0 = 0; s.z0 = 0; }
		scatter_hits3<MORTON>(hit, s.x0, s.y0, s.z0, g, table);     // converged: every write is predicated on its own bits
		return;
	}
	if (hit) scatter_hits3<MORTON>(hit, s.x0, s.y0, s.z0, g, table);
}

template <bool MORTON, bool SOA4>
__global__ void VOXB_TRI_BOUNDS surface_tri_kernel(const GridParams g, const float* __restrict__ tris,
                                                   unsigned int* __restrict__ table, const QueueView q) {
	__shared__ __align__(16) float stage[SOA4 ? 4 : (kTriBlock / 32) * 288];
	grid_launch_dependents();
	grid_dependency_wait();                  // the zero-fill / counter reset (or the tile kernel of a prepared mesh) in front of this kernel
	const unsigned long long tile = ((unsigned long long)blockIdx.x * kTriBlock + threadIdx.x) >> 5;
	tri_tile<MORTON, SOA4>(g, tris, table, q, tile, stage + (SOA4 ? 0 : (threadIdx.x >> 5) * 288));
}

// ------------------------------------------------------------------------------------------------
// Exact row solver.  For a fixed (y,z) row every x-dependent test value is
//     f(x) = fl(fl(c * fl(float(x) * unit.x)) + row constant ...)
// i.e. a composition of order-preserving maps of x (int->float conversion, rounding, multiplication by a
// constant, addition of a constant), hence weakly monotone in x with the direction given by the sign of c.
// So each test accepts a half-line of x (an interval for the plane test), and the exact end points follow from
// bisection with the reference expression itself as the predicate — no error analysis, no tolerance.
// ------------------------------------------------------------------------------------------------
// smallest x in [lo, hi] with pred(x) true, or hi + 1; pred must be false...false true...true on [lo, hi]
template <typename Pred>
__device__ __forceinline__ int first_true(int lo, int hi, Pred pred) {
	int L = lo, H = hi + 1;
	while (L < H) {
		const int mid = (L + H) >> 1;
		if (pred(mid)) H = mid; else L = mid + 1;
	}
	return L;
}
// largest x in [lo, hi] with pred(x) true, or lo - 1; pred must be true...true false...false on [lo, hi]
template <typename Pred>
__device__ __forceinline__ int last_true(int lo, int hi, Pred pred) {
	return first_true(lo, hi, [&](int x) { return !pred(x); }) - 1;
}
// The same two searches, started from an estimate `e` of the answer (the real-arithmetic root of the test, computed
// with a fast reciprocal).  The estimate only chooses WHERE the exact predicate is evaluated first: two evaluations
// confirm it when it is right (the common case), otherwise bisection finishes on the side the evaluations point to.
template <typename Pred>
__device__ __forceinline__ int first_true_from(int lo, int hi, int e, Pred pred) {
	e = max(lo, min(e, hi));
	if (pred(e)) {
		if (e == lo || !pred(e - 1)) return e;
		return first_true(lo, e - 1, pred);
	}
	if (e == hi) return hi + 1;
	if (pred(e + 1)) return e + 1;
	return first_true(e + 2, hi, pred);
}
template <typename Pred>
__device__ __forceinline__ int last_true_from(int lo, int hi, int e, Pred pred) {
	return first_true_from(lo, hi, e + 1, [&](int x) { return !pred(x); }) - 1;
}
// float -> int for estimates: saturating, NaN -> 0 (any value is acceptable, the predicate decides)
__device__ __forceinline__ int est_ceil(float v) { return __float2int_ru(v); }
__device__ __forceinline__ int est_floor(float v) { return __float2int_rd(v); }

// Narrow [lo, hi] to the x accepted by `acc`, which is monotone in x with direction sign(coef):
// coef > 0: rejected...accepted; coef < 0: accepted...rejected; otherwise constant in x.
// `root` estimates the x where the test value crosses zero.
template <typename Acc>
__device__ __forceinline__ void narrow(int& lo, int& hi, float coef, float root, Acc acc) {
	if (lo > hi) return;
	if (coef > 0.0f) lo = first_true_from(lo, hi, est_ceil(root), acc);
	else if (coef < 0.0f) hi = last_true_from(lo, hi, est_floor(root), acc);
	else if (!acc(lo)) hi = lo - 1;
}

// Accepted x-interval [lo, hi] of row (y,z) (empty when lo > hi); s.x0..s.x1 is the triangle's clipped bbox.
__device__ __forceinline__ void surf_solve_row(const SurfSetup& s, const GridParams& g, const SurfRow& r, int& lo, int& hi) {
	lo = s.x0; hi = s.x1;
	// plane (cpu_voxelizer.cpp:139-140): rejected iff P = (n.p + d1)(n.p + d2) > 0, i.e. both factors strictly on the
	// same side (and the product not underflowing to 0).  n.p moves with sign(n.x): the "both below" rejections are
	// closed towards one end of the row and the "both above" ones towards the other.
	auto plane = [&](int x, bool& below) {
		const float px = fmul((float)x, g.ux);
		const float ndp = fadd(fadd(fmul(s.nx, px), r.ny_py), r.nz_pz);
		const float a = fadd(ndp, s.d1);
		below = a < 0.0f;
		return fmul(a, fadd(ndp, s.d2)) > 0.0f;       // true = rejected
	};
	if (s.nx > 0.0f || s.nx < 0.0f) {
		// the two factors cross zero near xA and xB; the accepted voxels lie between them
		const float inv = __fdividef(1.0f, s.nx * g.ux), c = r.ny_py + r.nz_pz;
		const float xA = -(c + s.d1) * inv, xB = -(c + s.d2) * inv;
		const int e_lo = est_ceil(fminf(xA, xB)), e_hi = est_floor(fmaxf(xA, xB));
		if (s.nx > 0.0f) {
			lo = first_true_from(lo, hi, e_lo, [&](int x) { bool b; return !(plane(x, b) && b); });
			if (lo <= hi) hi = last_true_from(lo, hi, e_hi, [&](int x) { bool b; return !(plane(x, b) && !b); });
		} else {
			lo = first_true_from(lo, hi, e_lo, [&](int x) { bool b; return !(plane(x, b) && !b); });
			if (lo <= hi) hi = last_true_from(lo, hi, e_hi, [&](int x) { bool b; return !(plane(x, b) && b); });
		}
	} else {
		bool b;
		if (plane(lo, b)) hi = lo - 1;                // n.x is 0 or NaN: the test does not depend on x
	}
#pragma unroll
	for (int k = 0; k < 3; k++) {                      // XY edges (:144-147): value moves with sign(n_xy_e.x)
		const float root = __fdividef(-(r.xy_bpy[k] + s.xy_d[k]), s.xy_a[k] * g.ux);
		narrow(lo, hi, s.xy_a[k], root, [&](int x) { return !(fadd(fadd(fmul(s.xy_a[k], fmul((float)x, g.ux)), r.xy_bpy[k]), s.xy_d[k]) < 0.0f); });
	}
#pragma unroll
	for (int k = 0; k < 3; k++) {                      // ZX edges (:156-159): value moves with sign(n_zx_e.y)
		const float root = __fdividef(-(r.zx_apz[k] + s.zx_d[k]), s.zx_b[k] * g.ux);
		narrow(lo, hi, s.zx_b[k], root, [&](int x) { return !(fadd(fadd(r.zx_apz[k], fmul(s.zx_b[k], fmul((float)x, g.ux))), s.zx_d[k]) < 0.0f); });
	}
}

// Writes the accepted interval [xa, xb] of row (y,z) into the table as word masks.
template <bool MORTON>
__device__ __forceinline__ void write_row_interval(const GridParams& g, unsigned int* __restrict__ table, int y, int z, int xa, int xb) {
	if (MORTON) {
		// the row in morton order: y and z are spread once, x advances in dilated form (x + 1 on bits 0,3,6,...)
		const unsigned long long kX = 0x1249249249249249ull;
		const unsigned long long yz = (spread3((unsigned)y) << 1) | (spread3((unsigned)z) << 2);
		unsigned long long mx = spread3((unsigned)xa);
		WordRun<false> run;
		for (int x = xa; x <= xb; x++) {
			run.add(table, g, mx | yz);
			mx = ((mx | ~kX) + 1ull) & kX;
		}
		run.flush(table);
	} else if ((g.G & 31) != 0) {
		WordRun<false> run;
		for (int x = xa; x <= xb; x++) run.add(table, g, voxel_index<false>(g, x, y, z));
		run.flush(table);
	} else {
		// rows are whole words: first/last word get partial masks (x at bit 31 - x%32), the middle ones 0xffffffff
		unsigned int* rowp = table + (((unsigned long long)g.G * ((unsigned long long)y + (unsigned long long)g.G * (unsigned long long)z)) >> 5) - g.word_base;
		const int wa = xa >> 5, wb = xb >> 5;
		const unsigned int first = 0xffffffffu >> (xa & 31), last = 0xffffffffu << (31 - (xb & 31));
		if (wa == wb) {
			atomicOr(rowp + wa, first & last);
		} else {
			atomicOr(rowp + wa, first);
			for (int w = wa + 1; w < wb; w++) atomicOr(rowp + w, 0xffffffffu);
			atomicOr(rowp + wb, last);
		}
	}
}

// The queued triangle of slot `slot`: its stored setup, or (beyond the setup buffer) the setup recomputed from the triangle.
template <bool SOA4>
__device__ __forceinline__ void coop_setup(const GridParams& g, const float* __restrict__ tris, const QueueView& q, unsigned int slot, unsigned int tri, SurfSetup& s) {
	if (slot < q.setup_cap) {
		load_setup(q.setups + (size_t)slot * kSetupVec, s);       // computed once by the per-triangle kernel
	} else {
		Tri t;
		if (SOA4) load_tri_soa4(tris, g.n_tris, tri, t); else load_tri_aos(tris, tri, t);
		shift_tri(t, g);
		surf_setup(t, g, s);
		clip_to_region(g, s);
	}
}

// Persistent grid over the queued (y,z) rows: every queued triangle owns a run of consecutive row numbers (reserved
// together with its queue slot, so the queue is sorted by first row and a directory maps every 64th row to its slot).
// A warp takes 64 consecutive rows per iteration, whatever triangles they belong to:
//   phase 1  every lane finds the triangle of two rows and runs the x-independent YZ tests on them (4 of the setup's 10
//            vectors); the (slot, row) pairs that pass are compacted through shared memory;
//   phase 2  one surviving row per lane: full setup, exact row solver, word-mask writes.
// No lane sits out the long solve because its row failed the short test, for big triangles (all 64 rows of one
// triangle: the setup loads are broadcasts) and small ones (several triangles per warp) alike.
template <bool MORTON, bool SOA4>
__global__ void __launch_bounds__(kBlock) surface_coop_kernel(const GridParams g, const float* __restrict__ tris,
                                                              unsigned int* __restrict__ table,
                                                              const QueueView q) {
	grid_dependency_wait();                  // the queue is the per-triangle kernel's output
	const unsigned long long packed = *q.cursor;
	const unsigned int n_entries = (unsigned int)(packed >> 32);
	const unsigned int n_rows = (unsigned int)packed;
	__shared__ uint2 survivors[kBlock / 32][kRowsPerWarp];          // {slot, row within the triangle}
	const int lane = threadIdx.x & 31;
	uint2* mine = survivors[threadIdx.x >> 5];
	const unsigned int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
	const unsigned int n_warps = (gridDim.x * kBlock) >> 5;
	const unsigned int n_buckets = (n_rows + kRowsPerWarp - 1) / kRowsPerWarp;
	for (unsigned int b = warp; b < n_buckets; b += n_warps) {
		unsigned int pass[2];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const unsigned int r = b * kRowsPerWarp + 32 * h + lane;
			bool ok = r < n_rows;
			unsigned int slot = 0u, local = 0u;
			if (ok) {
				slot = find_slot(q, r, n_entries, n_rows);
				const uint2 e = __ldg(&q.entries[slot]);
				local = r - e.y;
				if (slot < q.setup_cap) {
					// the YZ part of the stored setup: yz_a, yz_b, yz_d and the bbox (vectors 3, 4, 5, 8)
					const uint4* sp = q.setups + (size_t)slot * kSetupVec;
					const uint4 v3 = __ldg(sp + 3), v4 = __ldg(sp + 4), v5 = __ldg(sp + 5), v8 = __ldg(sp + 8);
					const int y0 = (int)(v8.y & 0xffffu), ny = (int)(v8.y >> 16) - y0 + 1, z0 = (int)(v8.z & 0xffffu);
					const int y = y0 + (int)(local % (unsigned int)ny), z = z0 + (int)(local / (unsigned int)ny);
					const float py = fmul((float)y, g.uy), pz = fmul((float)z, g.uz);
					ok = !(fadd(dot2(__uint_as_float(v3.z), __uint_as_float(v4.y), py, pz), __uint_as_float(v5.x)) < 0.0f) &&
					     !(fadd(dot2(__uint_as_float(v3.w), __uint_as_float(v4.z), py, pz), __uint_as_float(v5.y)) < 0.0f) &&
					     !(fadd(dot2(__uint_as_float(v4.x), __uint_as_float(v4.w), py, pz), __uint_as_float(v5.z)) < 0.0f);
				} else {
					SurfSetup s;
					coop_setup<SOA4>(g, tris, q, slot, e.x, s);
					const int ny = s.y1 - s.y0 + 1;
					ok = surf_row_passes_yz(s, g, s.y0 + (int)(local % (unsigned int)ny), s.z0 + (int)(local / (unsigned int)ny));
				}
			}
			pass[h] = __ballot_sync(0xffffffffu, ok);
			if (ok) mine[(h ? __popc(pass[0]) : 0) + __popc(pass[h] & ((1u << lane) - 1u))] = make_uint2(slot, local);
		}
		const int n = __popc(pass[0]) + __popc(pass[1]);
		__syncwarp();
		for (int k = lane; k < n; k += 32) {
			const uint2 sr = mine[k];
			SurfSetup s;
			coop_setup<SOA4>(g, tris, q, sr.x, sr.x < q.setup_cap ? 0u : __ldg(&q.entries[sr.x].x), s);
			const int ny = s.y1 - s.y0 + 1;
			const int z = s.z0 + (int)(sr.y / (unsigned int)ny), y = s.y0 + (int)(sr.y % (unsigned int)ny);
			SurfRow row;
			surf_row_values(s, g, y, z, row);
			if (!MORTON && (g.G & 31) == 0 && s.x1 - s.x0 < kSweepWidth) {
				// a narrow row (triangles of a few voxels): testing its voxels one by one costs less than locating interval ends
				unsigned int m = 0u;
#pragma unroll 1
				for (int x = s.x0; x <= s.x1; x++) m |= surf_voxel(s, g, row, x) ? 0x80000000u >> (x - s.x0) : 0u;      // x0 at bit 31, MSB first like the table
				unsigned int* rowp = table + (((unsigned long long)g.G * ((unsigned long long)y + (unsigned long long)g.G * (unsigned long long)z)) >> 5) - g.word_base + (s.x0 >> 5);
				const unsigned int sh = (unsigned int)s.x0 & 31u, hi = m >> sh, lo = __funnelshift_r(0u, m, sh);
				if (hi) atomicOr(rowp, hi);
				if (lo) atomicOr(rowp + 1, lo);
				continue;
			}
			int xa, xb;
			surf_solve_row(s, g, row, xa, xb);
			if (xa <= xb) write_row_interval<MORTON>(g, table, y, z, xa, xb);
		}
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
template <bool MORTON, bool SOA4>
static cudaError_t run_surface(Workspace& ws, const GridParams& g, const float* d_tris, unsigned int* d_table, cudaStream_t st) {
	const unsigned long long tiles = (g.n_tris + 31ull) / 32ull;
	const unsigned long long blocks = (tiles + (kTriBlock / 32) - 1) / (kTriBlock / 32);
	cudaError_t err = launch_dependent(surface_tri_kernel<MORTON, SOA4>, (unsigned)blocks, kTriBlock, st, g, d_tris, d_table, ws.view());
	g_launch_count++;
	if (err == cudaSuccess) err = cudaGetLastError();
	if (err != cudaSuccess) return err;
	prof_mark(ws, 2, st);
	static int per_sm = 0;
	if (per_sm == 0) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, surface_coop_kernel<MORTON, SOA4>, kBlock, 0);
	if (err != cudaSuccess) return err;
	if (per_sm < 1) per_sm = 1;
	err = launch_dependent(surface_coop_kernel<MORTON, SOA4>, (unsigned)(ws.sm_count * per_sm), kBlock, st, g, d_tris, d_table, ws.view());
	g_launch_count++;
	return err != cudaSuccess ? err : cudaGetLastError();
}

cudaError_t launch_surface(Workspace& ws, const GridParams& g, const float* d_tris, unsigned int* d_table,
                           size_t region_words, const LaunchOpts& o, cudaStream_t st) {
	cudaError_t err = ensure_queue(ws, (size_t)g.n_tris);
	if (err != cudaSuccess) return err;
	prof_mark(ws, 0, st);
	if (!o.accumulate) err = launch_zero(ws, d_table, region_words, st, true);       // also resets the counters
	else err = cudaMemsetAsync(ws.counters, 0, kNumCounters * sizeof(unsigned long long), st);
	if (err != cudaSuccess) return err;
	prof_mark(ws, 1, st);
	if (g.n_tris != 0) {
		if (o.morton) err = o.soa4 ? run_surface<true, true>(ws, g, d_tris, d_table, st) : run_surface<true, false>(ws, g, d_tris, d_table, st);
		else err = o.soa4 ? run_surface<false, true>(ws, g, d_tris, d_table, st) : run_surface<false, false>(ws, g, d_tris, d_table, st);
		if (err != cudaSuccess) return err;
	} else {
		prof_mark(ws, 2, st);
	}
	prof_mark(ws, 3, st);
	prof_mark(ws, 4, st);
	if (ws.prof_on) ws.prof_calls++;
	return cudaSuccess;
}

}  // namespace voxb
