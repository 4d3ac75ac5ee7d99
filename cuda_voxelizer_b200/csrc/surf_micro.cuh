// surf_micro.cuh — the branch-free small-triangle evaluation of the surface path (<=3x3x3 and <=4x4x4 candidate boxes)
// and the warp-private triangle fetch, shared by the per-triangle kernel (surface.cu) and the tile-owner kernel (tiles.cu).
#pragma once
#include "vox_exact.cuh"

#ifndef VOXB_DEBUG_RED_MODE
#define VOXB_DEBUG_RED_MODE 0
#endif
#ifndef VOXB_DEBUG_WINDOW_WORDS
#define VOXB_DEBUG_WINDOW_WORDS 0x400000u
#endif

namespace voxb {

// atomicOr whose result is not wanted, as a reduction (ptxas keeps a predicated atomicOr as ATOMG, which makes the warp wait for
// the returned value at exit)
__device__ __forceinline__ void red_or(unsigned int* p, unsigned int v) {
	asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int sign_in(unsigned int mask, float v) {       // mask = (mask << 1) | signbit(v)
	return __funnelshift_l(__float_as_uint(v), mask, 1);
}
// "some edge function of the cell is negative" (cpu_voxelizer.cpp:145-159: three `< 0.0f` tests): the sign bit of a | b | c (one
// LOP3; fminf(fminf(a, b), c) compiles to one FMNMX3, so this is the same instruction count, stated directly).  Equal to testing each value: an edge value is a sum `... + d_e` whose last
// term is never -0 (d_e itself ends in `+ max(0, .)`, and x + (+0) is never -0), so no value is -0; a NaN comes out of FADD as
// the canonical 0x7fffffff, sign 0, and `NaN < 0` is false as well.
__device__ __forceinline__ float any_negative3(float a, float b, float c) {
	return __uint_as_float(__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c));
}

// The nine sums a[i] + b[j] of a 3x3 cell block in five additions (four of them FADD2), and one more level `+ d`
// in five again: p[j] = cells (0,j),(1,j); q = cells (2,0),(2,1) (computed as b + a: IEEE addition commutes bit
// for bit); r = cell (2,2).
struct Cells3 { float2 p[3]; float2 q; float r; };
__device__ __forceinline__ Cells3 sum3(const float a[3], const float b[3]) {
	Cells3 c;
	const float2 a01 = make_float2(a[0], a[1]), b01 = make_float2(b[0], b[1]);
#pragma unroll
	for (int j = 0; j < 3; j++) c.p[j] = fadd2(a01, bc(b[j]));
	c.q = fadd2(b01, bc(a[2]));
	c.r = fadd(a[2], b[2]);
	return c;
}
__device__ __forceinline__ Cells3 plus3(const Cells3& c, float d) {
	Cells3 o;
#pragma unroll
	for (int j = 0; j < 3; j++) o.p[j] = fadd2(c.p[j], bc(d));
	o.q = fadd2(c.q, bc(d));
	o.r = fadd(c.r, d);
	return o;
}
__device__ __forceinline__ float cell3(const Cells3& c, int i, int j) {
	if (i < 2) return i == 0 ? c.p[j].x : c.p[j].y;
	if (j < 2) return j == 0 ? c.q.x : c.q.y;
	return c.r;
}
// The three edge functions of one projection plane over the 3x3 cells: v[e] cell (i,j) = (ea[e]*pa[i] + eb[e]*pb[j]) + ed[e]
__device__ __forceinline__ void edge_cells3(const float ea[3], const float eb[3], const float ed[3], const float pa[3], const float pb[3], Cells3 v[3]) {
#pragma unroll
	for (int e = 0; e < 3; e++) {
		float a[3], b[3];
#pragma unroll
		for (int i = 0; i < 3; i++) { a[i] = fmul(ea[e], pa[i]); b[i] = fmul(eb[e], pb[i]); }
		v[e] = plus3(sum3(a, b), ed[e]);
	}
}

__device__ __forceinline__ unsigned int surf_micro3(const SurfSetup& s, const GridParams& g) {
	float px[3], py[3], pz[3];
#pragma unroll
	for (int i = 0; i < 3; i++) {
		px[i] = fmul((float)(s.x0 + i), g.ux);
		py[i] = fmul((float)(s.y0 + i), g.uy);
		pz[i] = fmul((float)(s.z0 + i), g.uz);
	}
	// plane: ((n.x*p.x + n.y*p.y) + n.z*p.z), then (s + d1) * (s + d2) > 0 rejects
	float nxp[3], nyp[3], nzp[3];
#pragma unroll
	for (int i = 0; i < 3; i++) { nxp[i] = fmul(s.nx, px[i]); nyp[i] = fmul(s.ny, py[i]); nzp[i] = fmul(s.nz, pz[i]); }
	const Cells3 t = sum3(nxp, nyp);
	unsigned int rej = 0u;
#pragma unroll
	for (int k = 2; k >= 0; k--) {
		const Cells3 sdp = plus3(t, nzp[k]);
		const Cells3 A = plus3(sdp, s.d1), B = plus3(sdp, s.d2);
		Cells3 N;                                   // 0 - A*B: the products are scalar FMULs (see vox_exact.cuh)
#pragma unroll
		for (int j = 0; j < 3; j++) N.p[j] = fsub2(bc(0.0f), make_float2(fmul(A.p[j].x, B.p[j].x), fmul(A.p[j].y, B.p[j].y)));
		N.q = fsub2(bc(0.0f), make_float2(fmul(A.q.x, B.q.x), fmul(A.q.y, B.q.y)));
		N.r = fsub(0.0f, fmul(A.r, B.r));
#pragma unroll
		for (int j = 2; j >= 0; j--)
#pragma unroll
			for (int i = 0; i < 3; i++) rej = sign_in(rej, cell3(N, i, j));      // x ascending: x0 at the HIGH bit of its row
	}
	unsigned int rxy = 0u, ryz = 0u, rzx = 0u;
	{
		Cells3 v[3];                                // XY cells (i,j): bit i + 3j, replicated over k
		edge_cells3(s.xy_a, s.xy_b, s.xy_d, px, py, v);
#pragma unroll
		for (int j = 2; j >= 0; j--)
#pragma unroll
			for (int i = 0; i < 3; i++) rxy = sign_in(rxy, any_negative3(cell3(v[0], i, j), cell3(v[1], i, j), cell3(v[2], i, j)));
	}
	{
		Cells3 v[3];                                // YZ cells (j,k): value = (n.x*p.y + n.y*p.z) + d; bit j + 3k, replicated over i
		edge_cells3(s.yz_a, s.yz_b, s.yz_d, py, pz, v);
#pragma unroll
		for (int k = 2; k >= 0; k--)
#pragma unroll
			for (int j = 2; j >= 0; j--) ryz = sign_in(ryz, any_negative3(cell3(v[0], j, k), cell3(v[1], j, k), cell3(v[2], j, k)));
	}
	{
		Cells3 v[3];                                // ZX cells (k,i): value = (n.x*p.z + n.y*p.x) + d; bit i + 3k, replicated over j
		edge_cells3(s.zx_a, s.zx_b, s.zx_d, pz, px, v);
#pragma unroll
		for (int k = 2; k >= 0; k--)
#pragma unroll
			for (int i = 0; i < 3; i++) rzx = sign_in(rzx, any_negative3(cell3(v[0], k, i), cell3(v[1], k, i), cell3(v[2], k, i)));
	}
	// expand the 9-bit cell masks to the 27-bit voxel layout b = (2-i) + 3j + 9k
	const unsigned int xy27 = rxy * 0x40201u;                                           // copies at +0, +9, +18
	const unsigned int yz_s = (ryz & 0x1u) | ((ryz & 0x2u) << 2) | ((ryz & 0x4u) << 4) | ((ryz & 0x8u) << 6) | ((ryz & 0x10u) << 8) |
	                          ((ryz & 0x20u) << 10) | ((ryz & 0x40u) << 12) | ((ryz & 0x80u) << 14) | ((ryz & 0x100u) << 16);   // bit (j+3k) -> 3j+9k
	const unsigned int yz27 = yz_s * 7u;                                                // copies at +0, +1, +2
	const unsigned int zx_s = (rzx & 0x7u) | ((rzx & 0x38u) << 6) | ((rzx & 0x1c0u) << 12);                                     // bit (i+3k) -> i+9k
	const unsigned int zx27 = zx_s * 0x49u;                                             // copies at +0, +3, +6
	const int ex = s.x1 - s.x0, ey = s.y1 - s.y0, ez = s.z1 - s.z0;                      // 0..2
	const unsigned int valid = (((7u << (2 - ex)) & 7u) * 0x1249249u) & (((8u << (3 * ey)) - 1u) * 0x40201u) & ((512u << (9 * ez)) - 1u);
	return valid & ~(rej | xy27 | yz27 | zx27);
}

// The same for a bbox of at most 4x4x4 voxels (triangles up to ~3 voxels across, whatever their alignment).
// 64 candidates: bit (3-i) + 4j + 16k of the result = voxel (x0+i, y0+j, z0+k), built as four 16-bit z-slices.
// Used for the whole warp as soon as one of its triangles does not fit 3x3x3.
struct Cells4 { float2 p[4][2]; };                  // p[j][h] = cells (2h, j), (2h+1, j)
__device__ __forceinline__ Cells4 sum4(const float a[4], const float b[4]) {
	Cells4 c;
	const float2 a01 = make_float2(a[0], a[1]), a23 = make_float2(a[2], a[3]);
#pragma unroll
	for (int j = 0; j < 4; j++) { c.p[j][0] = fadd2(a01, bc(b[j])); c.p[j][1] = fadd2(a23, bc(b[j])); }
	return c;
}
__device__ __forceinline__ Cells4 plus4(const Cells4& c, float d) {
	Cells4 o;
#pragma unroll
	for (int j = 0; j < 4; j++) { o.p[j][0] = fadd2(c.p[j][0], bc(d)); o.p[j][1] = fadd2(c.p[j][1], bc(d)); }
	return o;
}
__device__ __forceinline__ float cell4(const Cells4& c, int i, int j) { return (i & 1) ? c.p[j][i >> 1].y : c.p[j][i >> 1].x; }
__device__ __forceinline__ void edge_cells4(const float ea[3], const float eb[3], const float ed[3], const float pa[4], const float pb[4], Cells4 v[3]) {
#pragma unroll
	for (int e = 0; e < 3; e++) {
		float a[4], b[4];
#pragma unroll
		for (int i = 0; i < 4; i++) { a[i] = fmul(ea[e], pa[i]); b[i] = fmul(eb[e], pb[i]); }
		v[e] = plus4(sum4(a, b), ed[e]);
	}
}

__device__ __forceinline__ unsigned long long surf_micro4(const SurfSetup& s, const GridParams& g) {
	float px[4], py[4], pz[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		px[i] = fmul((float)(s.x0 + i), g.ux);
		py[i] = fmul((float)(s.y0 + i), g.uy);
		pz[i] = fmul((float)(s.z0 + i), g.uz);
	}
	float nxp[4], nyp[4], nzp[4];
#pragma unroll
	for (int i = 0; i < 4; i++) { nxp[i] = fmul(s.nx, px[i]); nyp[i] = fmul(s.ny, py[i]); nzp[i] = fmul(s.nz, pz[i]); }
	const Cells4 t = sum4(nxp, nyp);
	unsigned int rej[4] = {0u, 0u, 0u, 0u};
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const Cells4 sdp = plus4(t, nzp[k]);
		const Cells4 A = plus4(sdp, s.d1), B = plus4(sdp, s.d2);
		Cells4 N;
#pragma unroll
		for (int j = 0; j < 4; j++)
#pragma unroll
			for (int h = 0; h < 2; h++) N.p[j][h] = fsub2(bc(0.0f), make_float2(fmul(A.p[j][h].x, B.p[j][h].x), fmul(A.p[j][h].y, B.p[j][h].y)));
#pragma unroll
		for (int j = 3; j >= 0; j--)
#pragma unroll
			for (int i = 0; i < 4; i++) rej[k] = sign_in(rej[k], cell4(N, i, j));
	}
	unsigned int rxy = 0u, ryz = 0u, rzx = 0u;          // 16 cells each: (i,j) -> (3-i)+4j, (j,k) -> j+4k, (k,i) -> (3-i)+4k
	{
		Cells4 v[3];
		edge_cells4(s.xy_a, s.xy_b, s.xy_d, px, py, v);
#pragma unroll
		for (int j = 3; j >= 0; j--)
#pragma unroll
			for (int i = 0; i < 4; i++) rxy = sign_in(rxy, any_negative3(cell4(v[0], i, j), cell4(v[1], i, j), cell4(v[2], i, j)));
	}
	{
		Cells4 v[3];
		edge_cells4(s.yz_a, s.yz_b, s.yz_d, py, pz, v);
#pragma unroll
		for (int k = 3; k >= 0; k--)
#pragma unroll
			for (int j = 3; j >= 0; j--) ryz = sign_in(ryz, any_negative3(cell4(v[0], j, k), cell4(v[1], j, k), cell4(v[2], j, k)));
	}
	{
		Cells4 v[3];
		edge_cells4(s.zx_a, s.zx_b, s.zx_d, pz, px, v);
#pragma unroll
		for (int k = 3; k >= 0; k--)
#pragma unroll
			for (int i = 0; i < 4; i++) rzx = sign_in(rzx, any_negative3(cell4(v[0], k, i), cell4(v[1], k, i), cell4(v[2], k, i)));
	}
	const int ex = s.x1 - s.x0, ey = s.y1 - s.y0, ez = s.z1 - s.z0;                      // 0..3
	const unsigned int valid_xy = (((0xfu << (3 - ex)) & 0xfu) * 0x1111u) & ((16u << (4 * ey)) - 1u);
	unsigned long long hit = 0ull;
#pragma unroll
	for (int k = 3; k >= 0; k--) {
		const unsigned int m = (ryz >> (4 * k)) & 0xfu;                                  // row bits j of slice k -> 4 bits each
		const unsigned int yz16 = ((m & 1u) | ((m & 2u) << 3) | ((m & 4u) << 6) | ((m & 8u) << 9)) * 0xfu;
		const unsigned int zx16 = ((rzx >> (4 * k)) & 0xfu) * 0x1111u;                   // x bits of slice k, copied to every row
		const unsigned int h = (k <= ez) ? (valid_xy & ~(rej[k] | rxy | yz16 | zx16)) : 0u;
		hit = (hit << 16) | (unsigned long long)(h & 0xffffu);
	}
	return hit;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
	const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Fetches triangle `lane` of `tile` (32 consecutive triangles): the tile's 1152 bytes come in with 72 x 16-byte cp.async
// into the warp's shared slab and are read back at stride 9 words (conflict-free); scalar loads for the last,
// partial tile or an unaligned soup.
template <bool SOA4>
__device__ __forceinline__ bool load_tile_tri(const GridParams& g, const float* __restrict__ tris, unsigned long long tile, int lane,
                                              float* my_stage, Tri& t) {
	const unsigned long long i = (tile << 5) + lane;
	const bool valid = i < g.n_tris;
	const bool vec_ok = !SOA4 && (tile << 5) + 32ull <= g.n_tris && (reinterpret_cast<uintptr_t>(tris) & 15u) == 0;
	if (SOA4) {
		if (valid) load_tri_soa4(tris, g.n_tris, i, t);
	} else if (vec_ok) {
		const float4* src = reinterpret_cast<const float4*>(tris) + tile * 72ull;
		float4* dst = reinterpret_cast<float4*>(my_stage);
		cp_async16(dst + lane, src + lane);
		cp_async16(dst + 32 + lane, src + 32 + lane);
		if (lane < 8) cp_async16(dst + 64 + lane, src + 64 + lane);
		cp_async_wait_all();
		__syncwarp();
		const float* p = my_stage + 9 * lane;
		t.v0x = p[0]; t.v0y = p[1]; t.v0z = p[2];
		t.v1x = p[3]; t.v1y = p[4]; t.v1z = p[5];
		t.v2x = p[6]; t.v2y = p[7]; t.v2z = p[8];
	} else if (valid) {
		load_tri_aos(tris, i, t);
	}
	return valid;
}

// Writes a 64-bit hit mask of surf_micro4 into the table, one (y,z) row — four x-adjacent bits — at a time.
template <bool MORTON>
__device__ __forceinline__ void scatter_hits4(unsigned long long hit, int x0, int y0, int z0, const GridParams& g,
                                              unsigned int* __restrict__ table) {
	const bool fast = !MORTON && (g.G & 31) == 0 && g.w32;
	const unsigned int Gw = (unsigned int)g.G >> 5;
	const unsigned int w0 = fast ? Gw * ((unsigned int)y0 + (unsigned int)g.G * (unsigned int)z0) + ((unsigned int)x0 >> 5) - (unsigned int)g.word_base : 0u;
	const unsigned int sh = (unsigned int)x0 & 31u;
	while (hit) {
		const int r = (__ffsll((long long)hit) - 1) >> 2;      // row = j + 4k
		const unsigned int bits = (unsigned int)(hit >> (4 * r)) & 0xfu;        // x0 at bit 3 ... x0+3 at bit 0
		hit &= ~(0xfull << (4 * r));
		const int k = r >> 2, j = r & 3;
		if (fast) {
			const unsigned int w = w0 + Gw * ((unsigned int)j + (unsigned int)g.G * (unsigned int)k);
			const unsigned int v = bits << 28;
			const unsigned int hi = v >> sh, lo = __funnelshift_r(0u, v, sh);
			if (hi) atomicOr(table + w, hi);
			if (lo) atomicOr(table + w + 1, lo);
		} else {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				if (!((bits >> (3 - i)) & 1u)) continue;
				const unsigned long long idx = voxel_index<MORTON>(g, x0 + i, y0 + j, z0 + k);
				atomicOr(table + ((idx >> 5) - g.word_base), 1u << (31u - (unsigned int)(idx & 31ull)));
			}
		}
	}
}

// Writes a 27-bit hit mask (bit (2-i) + 3j + 9k = voxel (x0+i, y0+j, z0+k)) into the table: one (y,z) row — three
// x-adjacent bits — at a time, as one atomicOr, or two when the row straddles a word.
template <bool MORTON>
__device__ __forceinline__ void scatter_hits3(unsigned int hit, int x0, int y0, int z0, const GridParams& g,
                                              unsigned int* __restrict__ table) {
	if (MORTON) {
		unsigned long long cur = ~0ull;
		unsigned int mask = 0u;
		while (hit) {
			const int b = __ffs(hit) - 1;
			hit &= hit - 1u;
			const int k = b / 9, j = (b - 9 * k) / 3, i = 2 - (b - 9 * k - 3 * j);
			const unsigned long long idx = morton3((unsigned)(x0 + i), (unsigned)(y0 + j), (unsigned)(z0 + k));
			const unsigned long long w = (idx >> 5) - g.word_base;
			if (w != cur) { if (mask) atomicOr(table + cur, mask); cur = w; mask = 0u; }
			mask |= 1u << (31u - (unsigned int)(idx & 31ull));
		}
		if (mask) atomicOr(table + cur, mask);
		return;
	}
	if ((g.G & 31) == 0 && g.w32) {
		// rows are whole words and every word offset INSIDE THE REGION fits 32 bits: all-integer-32 addressing.  The absolute word
		// index may not (8192^3: 2^34 words) — the products below wrap modulo 2^32, and so does word_base: the difference is exact.
		const unsigned int Gw = (unsigned int)g.G >> 5;
		const unsigned int w0 = Gw * ((unsigned int)y0 + (unsigned int)g.G * (unsigned int)z0) + ((unsigned int)x0 >> 5) - (unsigned int)g.word_base;
		const unsigned int sh = (unsigned int)x0 & 31u;
		// straight-line: the nine rows one after the other, each write predicated on its bits (no loop, no divergence)
		unsigned int* p = table + w0;
		const unsigned int layer = Gw * (unsigned int)g.G;
#pragma unroll
		for (int k = 0; k < 3; k++)
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const int r = j + 3 * k;
				const unsigned int v = (hit << (29 - 3 * r)) & 0xe0000000u;       // the row's bits at the top: x0 -> bit 31
				const unsigned int hi = v >> sh, lo = __funnelshift_r(0u, v, sh);
#if VOXB_DEBUG_RED_MODE == 2      // experiment: all atomics folded into a 16 MB (L2-resident) window
				unsigned int* q = table + ((w0 + (unsigned int)j * Gw + (unsigned int)k * layer) & (VOXB_DEBUG_WINDOW_WORDS - 2u));
#else
				unsigned int* q = p + ((unsigned int)j * Gw + (unsigned int)k * layer);
#endif
#if VOXB_DEBUG_RED_MODE == 1      // experiment: no atomics (the compiler cannot prove the conditions false)
				if (hi == 0x12345678u) atomicOr(q, hi);
				if (lo == 0x12345678u) atomicOr(q + 1, lo);
#else
				if (hi) red_or(q, hi);
				if (lo) red_or(q + 1, lo);
#endif
			}
		return;
	}
	const unsigned long long G = (unsigned long long)g.G;
	while (hit) {
		const int r = (__ffs(hit) - 1) / 3;
		const unsigned int bits = (hit >> (3 * r)) & 7u;
		hit &= ~(7u << (3 * r));
		const int k = r / 3, j = r - 3 * k;
		const unsigned long long idx = (unsigned long long)x0 + G * ((unsigned long long)(y0 + j) + G * (unsigned long long)(z0 + k));
		// voxel idx+t sits at bit 31-((idx+t)&31): put the row MSB-first into a 64-bit window over words w, w+1
		const unsigned long long win = ((unsigned long long)bits << 61) >> (unsigned int)(idx & 31ull);
		const unsigned long long w = (idx >> 5) - g.word_base;
		const unsigned int hi = (unsigned int)(win >> 32), lo = (unsigned int)win;
		if (hi) atomicOr(table + w, hi);
		if (lo) atomicOr(table + w + 1, lo);
	}
}

}  // namespace voxb
