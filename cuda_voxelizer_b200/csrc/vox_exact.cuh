// vox_exact.cuh — the arithmetic of the hot path, written so the device reproduces the HOST
// float semantics of the reference's CPU voxelizer (the parity target):
//
//   * every binary32 operation is an explicitly rounded intrinsic (__fmul_rn/__fadd_rn/…): nvcc
//     never contracts those into FMA, whatever -fmad says;
//   * normalize() is IEEE sqrt then IEEE divide, not MUFU.RSQ (helper_math.h:78-81; §A-1);
//   * evaluation order follows cpu_voxelizer.cpp expression by expression (file:line cited below).
//
// Nothing here is a re-association: optimisations elsewhere only *skip* voxels that provably fail
// or hoist bit-identical sub-expressions out of loops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace voxb {

// ---------------------------------------------------------------------------------------------
// Launch-invariant grid parameters (from voxb200_grid / the reference's voxinfo, util.h:50-69)
// ---------------------------------------------------------------------------------------------
struct GridParams {
	float bx, by, bz;          // bbox.min
	float ux, uy, uz;          // unit
	float rux, ruy, ruz;       // fl(1 / unit): fast path of grid_coord() only, never part of a result
	int G;                     // gridsize (cubic, main.cpp:186)
	int rx0, rx1;              // region [lo, hi) per axis (z-slab, or morton octant)
	int ry0, ry1;
	int rz0, rz1;
	unsigned long long word_base;   // index of the table word holding the region's first voxel
	unsigned long long n_tris;
	int w32;                   // the region's words fit 32 bits: region-relative word offsets may be computed modulo 2^32
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// Two independent binary32 additions per instruction (FADD2, sm_100): each half is rounded to nearest exactly like
// __fadd_rn, so pairing additions changes no result bit.  ONLY additions are paired: ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, whereas it never fuses a scalar FMUL into a
// packed add — products therefore stay scalar (tests/test_abi.py checks the SASS holds no FFMA2 / FMUL2).
// A scalar broadcast operand (bc) costs nothing: FADD2 takes `R.F32` as its second source.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }   // a - b == a + (-b), bit for bit
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
// host fallbacks of helper_math.h:58-66 (a<b?a:b), and std::max(0.0f,x) (§A-8)
__device__ __forceinline__ float hmin(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float hmax(float a, float b) { return a > b ? a : b; }
// std::max(0.0f, x) = (0 < x) ? x : 0: equals fmaxf(x, +0) for EVERY input (NaN -> 0, -0 -> +0, x <= 0 -> +0), one FMNMX
__device__ __forceinline__ float max0(float x) { return fmaxf(x, 0.0f); }
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) {       // helper_math.h:1260
	return fadd(fmul(ax, bx), fmul(ay, by));
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {  // :1264
	return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }   // :1172

struct Tri {
	float v0x, v0y, v0z, v1x, v1y, v1z, v2x, v2y, v2z;    // already shifted by -bbox.min
};

// cpu_voxelizer.cpp:40-45 / voxelize.cu:71-73 — the vertex shift, one rounded subtraction per coordinate
__device__ __forceinline__ void shift_tri(Tri& t, const GridParams& g) {
	const float2 bxy = make_float2(g.bx, g.by);
	const float2 a = fsub2(make_float2(t.v0x, t.v0y), bxy), b = fsub2(make_float2(t.v1x, t.v1y), bxy), c = fsub2(make_float2(t.v2x, t.v2y), bxy);
	const float2 z01 = fsub2(make_float2(t.v0z, t.v1z), bc(g.bz));
	t.v0x = a.x; t.v0y = a.y; t.v1x = b.x; t.v1y = b.y; t.v2x = c.x; t.v2y = c.y;
	t.v0z = z01.x; t.v1z = z01.y; t.v2z = fsub(t.v2z, g.bz);
}

// normalize(cross(e0, e1)) — cpu_voxelizer.cpp:72 with helper_math.h:1436 (cross), :1325 + :78 (normalize)
__device__ __forceinline__ void tri_normal(float e0x, float e0y, float e0z, float e1x, float e1y, float e1z,
                                           float& nx, float& ny, float& nz) {
	float cx = fsub(fmul(e0y, e1z), fmul(e0z, e1y));
	float cy = fsub(fmul(e0z, e1x), fmul(e0x, e1z));
	float cz = fsub(fmul(e0x, e1y), fmul(e0y, e1x));
	float inv_len = fdiv(1.0f, __fsqrt_rn(dot3(cx, cy, cz, cx, cy, cz)));
	nx = fmul(cx, inv_len); ny = fmul(cy, inv_len); nz = fmul(cz, inv_len);
}

// ---------------------------------------------------------------------------------------------
// Surface: per-triangle constants of the Schwarz-Seidel overlap test (cpu_voxelizer.cpp:67-125)
// ---------------------------------------------------------------------------------------------
struct SurfSetup {
	float nx, ny, nz, d1, d2;
	float xy_a[3], xy_b[3], xy_d[3];    // n_xy_e*.x, n_xy_e*.y, d_xy_e*     (:91-101)
	float yz_a[3], yz_b[3], yz_d[3];    // n_yz_e*.x, n_yz_e*.y, d_yz_e*     (:103-113)
	float zx_a[3], zx_b[3], zx_d[3];    // n_zx_e*.x, n_zx_e*.y, d_xz_e*     (:115-125)
	int x0, x1, y0, y1, z0, z1;         // t_bbox_grid, clamped to the grid (:76-80) — NOT yet to the region
};

// One edge of one projection plane:
//   n_e = (-1*e.a, e.b), negated when the orthogonal normal component is < 0;
//   d_e = (-1*dot(n_e, v)) + max(0, ua*n_e.x) + max(0, ub*n_e.y)         (left to right)
__device__ __forceinline__ void edge_setup(float e_a, float e_b, bool flip, float va, float vb, float ua, float ub,
                                           float& na, float& nb, float& d) {
	na = fmul(-1.0f, e_a);
	nb = e_b;
	if (flip) { na = -na; nb = -nb; }
	d = fadd(fadd(fmul(-1.0f, dot2(na, nb, va, vb)), max0(fmul(ua, na))), max0(fmul(ub, nb)));
}

// :76-80 grid bbox = clamp(int(world / unit)).  Independent of everything else in the setup, so callers
// compute it first and drop triangles outside their region before paying for the rest.
// trunc(fl(v / u)) — the reference's float3_to_int3(world / unit) — without the IEEE division whenever the answer
// is unambiguous.  q = fl(v * fl(1/u)) differs from the correctly rounded quotient Q = fl(v/u) by less than
// |q| * 2^-22 (two roundings in q, one in Q).  If q is further than |q| * 2^-21 from both neighbouring integers then
// q and Q lie strictly inside the same (t, t+1) and truncate alike; otherwise (about one coordinate in 10^3) the
// real division decides.  Negative, huge and NaN quotients take the division too.
__device__ __forceinline__ int grid_coord(float v, float u, float ru) {
	const float q = fmul(v, ru);
	const int t = __float2int_rz(q);
	const float f = fsub(q, (float)t);
	const float tol = fmul(q, 4.76837158203125e-07f);        // |q| * 2^-21 (q > 0 on this path)
	if (q > 0.0f && q < 1.0e6f && f > tol && fsub(1.0f, f) > tol) return t;
	return __float2int_rz(fdiv(v, u));
}

__device__ __forceinline__ void surf_bbox(const Tri& t, const GridParams& g, SurfSetup& s) {
	const int gmax = g.G - 1;
	// min/max: the host fallbacks a<b?a:b / a>b?a:b (helper_math.h:58-66) equal FMNMX for all non-NaN inputs up to the
	// sign of a zero, which the int conversion below discards
	s.x0 = clampi(grid_coord(fminf(t.v0x, fminf(t.v1x, t.v2x)), g.ux, g.rux), 0, gmax);
	s.y0 = clampi(grid_coord(fminf(t.v0y, fminf(t.v1y, t.v2y)), g.uy, g.ruy), 0, gmax);
	s.z0 = clampi(grid_coord(fminf(t.v0z, fminf(t.v1z, t.v2z)), g.uz, g.ruz), 0, gmax);
	s.x1 = clampi(grid_coord(fmaxf(t.v0x, fmaxf(t.v1x, t.v2x)), g.ux, g.rux), 0, gmax);
	s.y1 = clampi(grid_coord(fmaxf(t.v0y, fmaxf(t.v1y, t.v2y)), g.uy, g.ruy), 0, gmax);
	s.z1 = clampi(grid_coord(fmaxf(t.v0z, fmaxf(t.v1z, t.v2z)), g.uz, g.ruz), 0, gmax);
}

// Everything but the bbox: normal, plane offsets, the 9 edge functions.
// The nine edge functions at once: products scalar, the 27 additions of the nine d_e as 15 (12 of them FADD2).
// Edge q = 3*plane + e.  d = ((-dot(n_e, v)) + max(0, ua*n_e.x)) + max(0, ub*n_e.y), left to right as in edge_setup;
// -1*x is the exact negation of x.
__device__ __forceinline__ void edge_setup9(const float e_a[9], const float e_b[9], const bool flip[3], const float va[9], const float vb[9],
                                            const float ua[3], const float ub[3], float na[9], float nb[9], float d[9]) {
	float m1[9], m2[9], t1[9], t2[9];
#pragma unroll
	for (int q = 0; q < 9; q++) {
		na[q] = fmul(-1.0f, e_a[q]);
		nb[q] = e_b[q];
		if (flip[q / 3]) { na[q] = -na[q]; nb[q] = -nb[q]; }
		m1[q] = fmul(na[q], va[q]); m2[q] = fmul(nb[q], vb[q]);
		t1[q] = max0(fmul(ua[q / 3], na[q])); t2[q] = max0(fmul(ub[q / 3], nb[q]));
	}
#pragma unroll
	for (int q = 0; q < 8; q += 2) {
		const float2 dot = fadd2(make_float2(m1[q], m1[q + 1]), make_float2(m2[q], m2[q + 1]));
		const float2 r = fadd2(fadd2(make_float2(-dot.x, -dot.y), make_float2(t1[q], t1[q + 1])), make_float2(t2[q], t2[q + 1]));
		d[q] = r.x; d[q + 1] = r.y;
	}
	d[8] = fadd(fadd(-fadd(m1[8], m2[8]), t1[8]), t2[8]);
}

// PLANE selects what is computed here: the normal and the plane offsets d1, d2 (:72, :83-87) — a prepared mesh stores them with
// the triangle and passes PLANE = false — and, always, the nine edge functions (:91-125).
template <bool PLANE = true>
__device__ __forceinline__ void surf_setup_tests(const Tri& t, const GridParams& g, SurfSetup& s) {
	// :68-70 edges
	const float2 v0 = make_float2(t.v0x, t.v0y), v1 = make_float2(t.v1x, t.v1y), v2 = make_float2(t.v2x, t.v2y);
	const float2 e0 = fsub2(v1, v0), e1 = fsub2(v2, v1), e2 = fsub2(v0, v2);
	const float2 ez01 = fsub2(make_float2(t.v1z, t.v2z), make_float2(t.v0z, t.v1z));
	const float e0x = e0.x, e0y = e0.y, e0z = ez01.x, e1x = e1.x, e1y = e1.y, e1z = ez01.y, e2x = e2.x, e2y = e2.y, e2z = fsub(t.v0z, t.v2z);
	if (PLANE) {
		tri_normal(e0x, e0y, e0z, e1x, e1y, e1z, s.nx, s.ny, s.nz);
		// :83-87 plane offsets
		float cx = (s.nx > 0.0f) ? g.ux : 0.0f;
		float cy = (s.ny > 0.0f) ? g.uy : 0.0f;
		float cz = (s.nz > 0.0f) ? g.uz : 0.0f;
		s.d1 = dot3(s.nx, s.ny, s.nz, fsub(cx, t.v0x), fsub(cy, t.v0y), fsub(cz, t.v0z));
		s.d2 = dot3(s.nx, s.ny, s.nz, fsub(fsub(g.ux, cx), t.v0x), fsub(fsub(g.uy, cy), t.v0y), fsub(fsub(g.uz, cz), t.v0z));
	}
	// :91-101 XY (n_e = (-e.y, e.x), flipped if n.z < 0; offsets pair unit.x/unit.y); :103-113 YZ (n_e = (-e.z, e.y), flipped if
	// n.x < 0; unit.y/unit.z); :115-125 ZX (n_e = (-e.x, e.z), flipped if n.y < 0).  The reference pairs unit.X with the first
	// ZX component (the coefficient of p.z) and unit.Z with the second (§A-13): kept verbatim.
	const float e_a[9] = {e0y, e1y, e2y, e0z, e1z, e2z, e0x, e1x, e2x};
	const float e_b[9] = {e0x, e1x, e2x, e0y, e1y, e2y, e0z, e1z, e2z};
	const float va[9] = {t.v0x, t.v1x, t.v2x, t.v0y, t.v1y, t.v2y, t.v0z, t.v1z, t.v2z};
	const float vb[9] = {t.v0y, t.v1y, t.v2y, t.v0z, t.v1z, t.v2z, t.v0x, t.v1x, t.v2x};
	const bool flip[3] = {s.nz < 0.0f, s.nx < 0.0f, s.ny < 0.0f};
	const float ua[3] = {g.ux, g.uy, g.ux}, ub[3] = {g.uy, g.uz, g.uz};
	float na[9], nb[9], d[9];
	edge_setup9(e_a, e_b, flip, va, vb, ua, ub, na, nb, d);
#pragma unroll
	for (int e = 0; e < 3; e++) {
		s.xy_a[e] = na[e]; s.xy_b[e] = nb[e]; s.xy_d[e] = d[e];
		s.yz_a[e] = na[3 + e]; s.yz_b[e] = nb[3 + e]; s.yz_d[e] = d[3 + e];
		s.zx_a[e] = na[6 + e]; s.zx_b[e] = nb[6 + e]; s.zx_d[e] = d[6 + e];
	}
}

__device__ __forceinline__ void surf_setup(const Tri& t, const GridParams& g, SurfSetup& s) {
	surf_bbox(t, g, s);
	surf_setup_tests(t, g, s);
}

// A queued triangle's setup travels from the per-triangle kernel to the cooperative kernel as 10 x 16 bytes
// (the bbox is the region-clipped one; coordinates fit 16 bits for G <= 65536).
constexpr int kSetupVec = 10;
__device__ __forceinline__ void store_setup(uint4* __restrict__ dst, const SurfSetup& s) {
#define VOXB_U(x) __float_as_uint(x)
	dst[0] = make_uint4(VOXB_U(s.nx), VOXB_U(s.ny), VOXB_U(s.nz), VOXB_U(s.d1));
	dst[1] = make_uint4(VOXB_U(s.d2), VOXB_U(s.xy_a[0]), VOXB_U(s.xy_a[1]), VOXB_U(s.xy_a[2]));
	dst[2] = make_uint4(VOXB_U(s.xy_b[0]), VOXB_U(s.xy_b[1]), VOXB_U(s.xy_b[2]), VOXB_U(s.xy_d[0]));
	dst[3] = make_uint4(VOXB_U(s.xy_d[1]), VOXB_U(s.xy_d[2]), VOXB_U(s.yz_a[0]), VOXB_U(s.yz_a[1]));
	dst[4] = make_uint4(VOXB_U(s.yz_a[2]), VOXB_U(s.yz_b[0]), VOXB_U(s.yz_b[1]), VOXB_U(s.yz_b[2]));
	dst[5] = make_uint4(VOXB_U(s.yz_d[0]), VOXB_U(s.yz_d[1]), VOXB_U(s.yz_d[2]), VOXB_U(s.zx_a[0]));
	dst[6] = make_uint4(VOXB_U(s.zx_a[1]), VOXB_U(s.zx_a[2]), VOXB_U(s.zx_b[0]), VOXB_U(s.zx_b[1]));
	dst[7] = make_uint4(VOXB_U(s.zx_b[2]), VOXB_U(s.zx_d[0]), VOXB_U(s.zx_d[1]), VOXB_U(s.zx_d[2]));
	dst[8] = make_uint4((unsigned)s.x0 | ((unsigned)s.x1 << 16), (unsigned)s.y0 | ((unsigned)s.y1 << 16), (unsigned)s.z0 | ((unsigned)s.z1 << 16), 0u);
	dst[9] = make_uint4(0u, 0u, 0u, 0u);
#undef VOXB_U
}
__device__ __forceinline__ void load_setup(const uint4* __restrict__ src, SurfSetup& s) {
#define VOXB_F(x) __uint_as_float(x)
	uint4 v;
	v = __ldg(src + 0); s.nx = VOXB_F(v.x); s.ny = VOXB_F(v.y); s.nz = VOXB_F(v.z); s.d1 = VOXB_F(v.w);
	v = __ldg(src + 1); s.d2 = VOXB_F(v.x); s.xy_a[0] = VOXB_F(v.y); s.xy_a[1] = VOXB_F(v.z); s.xy_a[2] = VOXB_F(v.w);
	v = __ldg(src + 2); s.xy_b[0] = VOXB_F(v.x); s.xy_b[1] = VOXB_F(v.y); s.xy_b[2] = VOXB_F(v.z); s.xy_d[0] = VOXB_F(v.w);
	v = __ldg(src + 3); s.xy_d[1] = VOXB_F(v.x); s.xy_d[2] = VOXB_F(v.y); s.yz_a[0] = VOXB_F(v.z); s.yz_a[1] = VOXB_F(v.w);
	v = __ldg(src + 4); s.yz_a[2] = VOXB_F(v.x); s.yz_b[0] = VOXB_F(v.y); s.yz_b[1] = VOXB_F(v.z); s.yz_b[2] = VOXB_F(v.w);
	v = __ldg(src + 5); s.yz_d[0] = VOXB_F(v.x); s.yz_d[1] = VOXB_F(v.y); s.yz_d[2] = VOXB_F(v.z); s.zx_a[0] = VOXB_F(v.w);
	v = __ldg(src + 6); s.zx_a[1] = VOXB_F(v.x); s.zx_a[2] = VOXB_F(v.y); s.zx_b[0] = VOXB_F(v.z); s.zx_b[1] = VOXB_F(v.w);
	v = __ldg(src + 7); s.zx_b[2] = VOXB_F(v.x); s.zx_d[0] = VOXB_F(v.y); s.zx_d[1] = VOXB_F(v.z); s.zx_d[2] = VOXB_F(v.w);
	v = __ldg(src + 8);
	s.x0 = (int)(v.x & 0xffffu); s.x1 = (int)(v.x >> 16); s.y0 = (int)(v.y & 0xffffu); s.y1 = (int)(v.y >> 16); s.z0 = (int)(v.z & 0xffffu); s.z1 = (int)(v.z >> 16);
#undef VOXB_F
}

// Per-(y,z)-row values of the test (cpu_voxelizer.cpp:138-159 with the x-independent products
// hoisted; each hoisted value is the same rounded product the reference recomputes per voxel).
struct SurfRow {
	float ny_py, nz_pz;          // n.y*p.y, n.z*p.z
	float xy_bpy[3];             // n_xy_e.y * p.y
	float zx_apz[3];             // n_zx_e.x * p.z
};

// The three YZ edge tests (:150-153): x-independent, so they accept or reject a whole (y,z) row.
__device__ __forceinline__ bool surf_row_passes_yz(const SurfSetup& s, const GridParams& g, int y, int z) {
	const float py = fmul((float)y, g.uy), pz = fmul((float)z, g.uz);       // :138
#pragma unroll
	for (int k = 0; k < 3; k++)
		if (fadd(dot2(s.yz_a[k], s.yz_b[k], py, pz), s.yz_d[k]) < 0.0f) return false;
	return true;
}
// The hoisted per-row products of a row that passed.
__device__ __forceinline__ void surf_row_values(const SurfSetup& s, const GridParams& g, int y, int z, SurfRow& r) {
	const float py = fmul((float)y, g.uy), pz = fmul((float)z, g.uz);
	r.ny_py = fmul(s.ny, py);
	r.nz_pz = fmul(s.nz, pz);
#pragma unroll
	for (int k = 0; k < 3; k++) { r.xy_bpy[k] = fmul(s.xy_b[k], py); r.zx_apz[k] = fmul(s.zx_a[k], pz); }
}
// Returns false when the YZ tests reject the whole row.
__device__ __forceinline__ bool surf_row(const SurfSetup& s, const GridParams& g, int y, int z, SurfRow& r) {
	if (!surf_row_passes_yz(s, g, y, z)) return false;
	surf_row_values(s, g, y, z, r);
	return true;
}

// The x-dependent tests of one voxel: plane (:139-140), XY (:144-147), ZX (:156-159).
__device__ __forceinline__ bool surf_voxel(const SurfSetup& s, const GridParams& g, const SurfRow& r, int x) {
	const float px = fmul((float)x, g.ux);
	const float ndp = fadd(fadd(fmul(s.nx, px), r.ny_py), r.nz_pz);
	if (fmul(fadd(ndp, s.d1), fadd(ndp, s.d2)) > 0.0f) return false;
#pragma unroll
	for (int k = 0; k < 3; k++)
		if (fadd(fadd(fmul(s.xy_a[k], px), r.xy_bpy[k]), s.xy_d[k]) < 0.0f) return false;
#pragma unroll
	for (int k = 0; k < 3; k++)
		if (fadd(fadd(r.zx_apz[k], fmul(s.zx_b[k], px)), s.zx_d[k]) < 0.0f) return false;
	return true;
}

// ---------------------------------------------------------------------------------------------
// Solid: yz-projection point-in-triangle at voxel centres (cpu_voxelizer.cpp:196-238, 254-296)
// ---------------------------------------------------------------------------------------------
// `fabs(t) < 0.000001` compares against a DOUBLE constant (cpu_voxelizer.cpp:2).  For a float f,
// (double)f < 1e-6  <=>  f < (smallest float >= 1e-6): a float compare with the rounded-up constant.
__device__ __forceinline__ float solid_eps() { return __double2float_ru(0.000001); }

struct SolidSetup {
	float nx, ny, nz;
	float v0x, v0y, v0z;                 // 3D v0 (get_x_coordinate uses the un-swapped v0, :296)
	float ay, az, by, bz, cy, cz;        // yz-projected, CCW-ordered v0/v1/v2 (:268-278)
	bool tl0, tl1, tl2;                  // TopLeftEdge of (a,b), (b,c), (c,a)   (:196-198, :294)
	int y0, y1, z0, z1;                  // centre-sample bbox (:285-290), clamped to the grid
	bool skip;                           // fabs(n.x) < float_error (:265)
};

__device__ __forceinline__ bool top_left(float v0x, float v0y, float v1x, float v1y) {
	return (v1y < v0y) || (v1y == v0y && v0x > v1x);
}

// floor (UP = false) or ceil (UP = true) of fl(fl(v / u) - 0.5f) — the centre-sample box of :282-286 — without the IEEE division
// whenever the answer is unambiguous, like grid_coord(): q = fl(v * fl(1/u)) is within |q| * 2^-22 of the correctly rounded
// quotient Q; for 1 <= q < 2^22 both q - 0.5 and Q - 0.5 are exact, so when q - 0.5 is further than |q| * 2^-21 from the integers
// below and above it, Q - 0.5 lies strictly between the same two integers and floors / ceils alike (ceil = floor + 1 there).
// Everything else (about one coordinate in 10^3, and values below one voxel) takes the real division.
template <bool UP>
__device__ __forceinline__ float half_coord(float v, float u, float ru) {
	const float q = fmul(v, ru);
	const float h = fsub(q, 0.5f);
	const float fl = floorf(h);
	const float f = fsub(h, fl);
	const float tol = fmul(q, 4.76837158203125e-07f);        // |q| * 2^-21
	if (q >= 1.0f && q < 4.0e6f && f > tol && fsub(1.0f, f) > tol) return UP ? fadd(fl, 1.0f) : fl;
	const float e = fsub(fdiv(v, u), 0.5f);
	return UP ? ceilf(e) : floorf(e);
}

__device__ __forceinline__ void solid_setup(const Tri& t, const GridParams& g, SolidSetup& s) {
	float e0x = fsub(t.v1x, t.v0x), e0y = fsub(t.v1y, t.v0y), e0z = fsub(t.v1z, t.v0z);
	float e1x = fsub(t.v2x, t.v1x), e1y = fsub(t.v2y, t.v1y), e1z = fsub(t.v2z, t.v1z);
	tri_normal(e0x, e0y, e0z, e1x, e1y, e1z, s.nx, s.ny, s.nz);
	s.skip = fabsf(s.nx) < solid_eps();
	s.v0x = t.v0x; s.v0y = t.v0y; s.v0z = t.v0z;
	s.ay = t.v0y; s.az = t.v0z; s.by = t.v1y; s.bz = t.v1z; s.cy = t.v2y; s.cz = t.v2z;
	// checkCCW (:201-209): (v1-v0) x (v2-v0) > 0, else swap v1 and v2
	float f0x = fsub(s.by, s.ay), f0y = fsub(s.bz, s.az), f1x = fsub(s.cy, s.ay), f1y = fsub(s.cz, s.az);
	float ccw = fsub(fmul(f0x, f1y), fmul(f1x, f0y));
	if (!(ccw > 0.0f)) { float ty = s.by, tz = s.bz; s.by = s.cy; s.bz = s.cz; s.cy = ty; s.cz = tz; }
	s.tl0 = top_left(s.ay, s.az, s.by, s.bz);
	s.tl1 = top_left(s.by, s.bz, s.cy, s.cz);
	s.tl2 = top_left(s.cy, s.cz, s.ay, s.az);
	// :282-286 — floor(max/unit - 0.5f), ceil(min/unit - 0.5f), all binary32
	float mxy = hmax(s.ay, hmax(s.by, s.cy)), mxz = hmax(s.az, hmax(s.bz, s.cz));
	float mny = hmin(s.ay, hmin(s.by, s.cy)), mnz = hmin(s.az, hmin(s.bz, s.cz));
	float fy1 = half_coord<false>(mxy, g.uy, g.ruy), fz1 = half_coord<false>(mxz, g.uz, g.ruz);
	float fy0 = half_coord<true>(mny, g.uy, g.ruy), fz0 = half_coord<true>(mnz, g.uz, g.ruz);
	// The reference does not clamp these to the grid (it would write out of bounds); they are
	// inside [0, G-1] whenever voxinfo encloses the mesh.  Clamp: deviates only where the reference is UB.
	const int gmax = g.G - 1;
	s.y0 = max(__float2int_rz(fy0), 0); s.z0 = max(__float2int_rz(fz0), 0);
	s.y1 = min(__float2int_rz(fy1), gmax); s.z1 = min(__float2int_rz(fz1), gmax);
	if (!(fy0 <= fy1) || !(fz0 <= fz1)) { s.y1 = s.y0 - 1; }   // NaN / empty range: no samples
}

// check_point_triangle (:217-238) fused with the accept rule of :294.  true = flip this column.
__device__ __forceinline__ bool solid_sample(const SolidSetup& s, float py, float pz) {
	const float eps = solid_eps();
	float PAx = fsub(py, s.ay), PAy = fsub(pz, s.az);
	float PBx = fsub(py, s.by), PBy = fsub(pz, s.bz);
	float PCx = fsub(py, s.cy), PCy = fsub(pz, s.cz);
	float t1 = fsub(fmul(PAx, PBy), fmul(PAy, PBx));
	if (fabsf(t1) < eps && fmul(PAx, PBx) <= 0.0f && fmul(PAy, PBy) <= 0.0f) return s.tl0;     // checknum 1
	float t2 = fsub(fmul(PBx, PCy), fmul(PBy, PCx));
	if (fabsf(t2) < eps && fmul(PBx, PCx) <= 0.0f && fmul(PBy, PCy) <= 0.0f) return s.tl1;     // checknum 2
	float t3 = fsub(fmul(PCx, PAy), fmul(PCy, PAx));
	if (fabsf(t3) < eps && fmul(PCx, PAx) <= 0.0f && fmul(PCy, PAy) <= 0.0f) return s.tl2;     // checknum 3
	return (fmul(t1, t2) > 0.0f) && (fmul(t1, t3) > 0.0f);                                      // checknum 0 / -1
}

// :292 sample point, :212-214 get_x_coordinate, :296 xmax = int(x / unit.x - 0.5) with a DOUBLE subtraction
__device__ __forceinline__ float solid_center(int i, float unit) { return fmul(fadd((float)i, 0.5f), unit); }
__device__ __forceinline__ int solid_xmax(const SolidSetup& s, const GridParams& g, float py, float pz) {
	float num = fadd(fmul(s.ny, fsub(py, s.v0y)), fmul(s.nz, fsub(pz, s.v0z)));
	float gx = fadd(fdiv(-num, s.nx), s.v0x);
	float q = fdiv(gx, g.ux);
	return __double2int_rz((double)q - 0.5);
}

// ---------------------------------------------------------------------------------------------
// Morton: LUT-free 3-way bit interleave in registers; equals the reference's LUT encoder
// (voxelize.cuh:20-34, morton_LUTs.h) for every coordinate below 2^16.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long spread3(unsigned int v) {
	unsigned long long x = v & 0x1fffffull;
	x = (x | (x << 32)) & 0x1f00000000ffffull;
	x = (x | (x << 16)) & 0x1f0000ff0000ffull;
	x = (x | (x << 8)) & 0x100f00f00f00f00full;
	x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
	x = (x | (x << 2)) & 0x1249249249249249ull;
	return x;
}
__host__ __device__ __forceinline__ unsigned long long morton3(unsigned int x, unsigned int y, unsigned int z) {
	return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// ---------------------------------------------------------------------------------------------
// Bit table (voxelize.cu:50-55 / util.h:25-38): word = idx/32, bit = 31 - idx%32 (x=0 at the MSB)
// ---------------------------------------------------------------------------------------------
template <bool MORTON>
__device__ __forceinline__ unsigned long long voxel_index(const GridParams& g, int x, int y, int z) {
	if (MORTON) return morton3((unsigned)x, (unsigned)y, (unsigned)z);
	return (unsigned long long)x + (unsigned long long)g.G * ((unsigned long long)y + (unsigned long long)g.G * (unsigned long long)z);
}

// Collects hits that fall in the same table word and issues ONE atomic per word run.
template <bool XOR>
struct WordRun {
	unsigned long long word;
	unsigned int mask;
	__device__ __forceinline__ WordRun() : word(~0ull), mask(0u) {}
	__device__ __forceinline__ void flush(unsigned int* table) {
		if (mask) { if (XOR) atomicXor(table + word, mask); else atomicOr(table + word, mask); }
		mask = 0u;
	}
	__device__ __forceinline__ void add(unsigned int* table, const GridParams& g, unsigned long long idx) {
		const unsigned long long w = (idx >> 5) - g.word_base;
		if (w != word) { flush(table); word = w; }
		const unsigned int bit = 1u << (31u - (unsigned int)(idx & 31ull));
		if (XOR) mask ^= bit; else mask |= bit;
	}
};

// ---------------------------------------------------------------------------------------------
// Triangle fetch.  AoS: the reference's 9-float records (main.cpp:61-80).  SoA4: 3 float4 planes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tri_aos(const float* __restrict__ tris, unsigned long long i, Tri& t) {
	const float* p = tris + 9ull * i;
	t.v0x = __ldg(p + 0); t.v0y = __ldg(p + 1); t.v0z = __ldg(p + 2);
	t.v1x = __ldg(p + 3); t.v1y = __ldg(p + 4); t.v1z = __ldg(p + 5);
	t.v2x = __ldg(p + 6); t.v2y = __ldg(p + 7); t.v2z = __ldg(p + 8);
}
__device__ __forceinline__ void load_tri_soa4(const float* __restrict__ tris, unsigned long long n, unsigned long long i, Tri& t) {
	const float4* p = reinterpret_cast<const float4*>(tris);
	const float4 a = __ldg(p + i), b = __ldg(p + n + i), c = __ldg(p + 2ull * n + i);
	t.v0x = a.x; t.v0y = a.y; t.v0z = a.z;
	t.v1x = b.x; t.v1y = b.y; t.v1z = b.z;
	t.v2x = c.x; t.v2y = c.y; t.v2z = c.z;
}

// Block-cooperative AoS fetch: the block's BLOCK triangles are one contiguous 36*BLOCK-byte run, read
// with 16-byte loads into shared memory, then each thread picks its 9 floats (stride 9 words:
// conflict-free).  Falls back to scalar loads when the base is not 16-byte aligned or at the tail.
template <int BLOCK>
__device__ __forceinline__ void load_tri_block_aos(const float* __restrict__ tris, unsigned long long n_tris,
                                                   unsigned long long block_first, float* smem, Tri& t, bool& valid) {
	const unsigned long long i = block_first + threadIdx.x;
	valid = i < n_tris;
	const bool full = (block_first + BLOCK <= n_tris) && ((reinterpret_cast<uintptr_t>(tris) & 15u) == 0);
	if (full) {
		const float4* src = reinterpret_cast<const float4*>(tris + 9ull * block_first);
		float4* dst = reinterpret_cast<float4*>(smem);
#pragma unroll
		for (int k = threadIdx.x; k < BLOCK * 9 / 4; k += BLOCK) dst[k] = __ldg(src + k);
		__syncthreads();
		const float* p = smem + 9 * threadIdx.x;
		t.v0x = p[0]; t.v0y = p[1]; t.v0z = p[2];
		t.v1x = p[3]; t.v1y = p[4]; t.v1z = p[5];
		t.v2x = p[6]; t.v2y = p[7]; t.v2z = p[8];
	} else if (valid) {
		load_tri_aos(tris, i, t);
	}
}

// ---------------------------------------------------------------------------------------------
// Work queue of the cooperative (large-triangle) kernels
// ---------------------------------------------------------------------------------------------
// Work units: (y,z) rows for the surface path, blocks of kSamplesPerItem centre samples for the solid path.
// Reserves, for every pushing lane of the warp, one queue slot and `items` consecutive work units with a
// single packed atomicAdd ((slots << 32) | items).  Because both halves advance together, queue[] ends
// up sorted by first-item, which is what lets the cooperative kernel binary-search item -> triangle.
// The cooperative-path work queue as the kernels see it.
struct QueueView {
	uint2* entries;                 // {triangle, first work item}, sorted by first item (slots and items are reserved together)
	unsigned long long* cursor;     // packed (entries << 32) | items
	uint4* setups;                  // stored SurfSetup of slot < setup_cap
	unsigned int setup_cap;
	unsigned int* dir;              // dir[b] = slot of the triangle that owns work item 64*b (b < dir_cap): bounds the item -> slot search
	unsigned int dir_cap;
};
constexpr int kDirShift = 6;

// Returns the caller's queue slot (0xffffffff when it did not push).
__device__ __forceinline__ unsigned int enqueue_warp(bool push, unsigned int items, unsigned int tri, const QueueView& q) {
	const unsigned int pushers = __ballot_sync(0xffffffffu, push);
	if (pushers == 0u) return 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned int mine = push ? items : 0u;
	unsigned int incl = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += up;
	}
	const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0ull;
	if (lane == 0) {
		base = atomicAdd(q.cursor, ((unsigned long long)__popc(pushers) << 32) | (unsigned long long)total);
		// the unit count lives in 32 bits: a call that queues more than 2^32 rows is flagged (cursor[3]) and reported by the host side
		if ((unsigned int)base + total < (unsigned int)base) atomicAdd(q.cursor + 3, 1ull);
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	const unsigned int slot = push ? (unsigned int)(base >> 32) + __popc(pushers & ((1u << lane) - 1u)) : 0xffffffffu;
	const unsigned int first = (unsigned int)base + (incl - mine);
	if (push) q.entries[slot] = make_uint2(tri, first);
	// every multiple of 64 inside [first, first + items) gets this slot in the directory: short ranges by their own lane,
	// long ones (a triangle of a million rows has 16 k entries) by the whole warp, one pusher after the other
	const unsigned int b0 = (first + (1u << kDirShift) - 1u) >> kDirShift;
	const unsigned long long end = (unsigned long long)first + mine;
	const unsigned int b1 = (unsigned int)min((unsigned long long)q.dir_cap, (end + (1ull << kDirShift) - 1ull) >> kDirShift);      // one past the last bucket
	const bool wide = push && b1 > b0 + 8u;
	if (push && !wide) for (unsigned int b = b0; b < b1; b++) q.dir[b] = slot;
	unsigned int todo = __ballot_sync(0xffffffffu, wide);
	while (todo) {
		const int src = __ffs(todo) - 1;
		todo &= todo - 1u;
		const unsigned int f0 = __shfl_sync(0xffffffffu, b0, src), f1 = __shfl_sync(0xffffffffu, b1, src), sl = __shfl_sync(0xffffffffu, slot, src);
		for (unsigned int b = f0 + lane; b < f1; b += 32u) q.dir[b] = sl;
	}
	return slot;
}

// Slot of the queued triangle that owns work item `item` (n_entries, n_items from the cursor).
__device__ __forceinline__ unsigned int find_slot(const QueueView& q, unsigned int item, unsigned int n_entries, unsigned int n_items) {
	const unsigned int b = item >> kDirShift;
	unsigned int lo = b < q.dir_cap ? __ldg(q.dir + b) : 0u;
	unsigned int hi = (b + 1u < q.dir_cap && ((b + 1u) << kDirShift) < n_items) ? __ldg(q.dir + b + 1u) : n_entries - 1u;
	while (lo < hi) {                                   // last entry whose first item <= item
		const unsigned int mid = (lo + hi + 1u) >> 1;
		if (__ldg(&q.entries[mid].y) <= item) lo = mid; else hi = mid - 1u;
	}
	return lo;
}

}  // namespace voxb
