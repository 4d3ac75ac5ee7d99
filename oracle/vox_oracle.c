/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle: a plain-C restatement of the voxelization hot path of Forceflow/cuda_voxelizer,
 * following the reference's *CPU* implementation (the parity target named by BASELINE.json):
 *
 *   surface  ....  /root/reference/src/cpu_voxelizer.cpp:35-176
 *   solid  ......  /root/reference/src/cpu_voxelizer.cpp:196-238 (helpers), 241-312
 *   morton  .....  /root/reference/src/cpu_voxelizer.cpp:18-32 + src/morton_LUTs.h
 *   bit layout ..  /root/reference/src/cpu_voxelizer.cpp:7-15, src/util.h:25-38
 *   bbox cube ...  /root/reference/src/util.h:80-110 ; unit: util.h:56-61
 *   vector math .  /root/reference/src/libs/cuda/helper_math.h:51-82 (host fallbacks),
 *                  1260-1267 (dot), 1325-1329 (normalize), 1436-1439 (cross)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (cuda_voxelizer_b200/) never links or calls it.
 *
 * PARITY PIN: the reference ships no golden vectors (its CI only checks an exit code), so this
 * restatement is pinned against outputs of the reference itself run here: oracle/_ref/libvoxref.so
 * is the reference's UNMODIFIED cpu_voxelizer.cpp compiled from /root/reference (oracle/Makefile),
 * tests/test_oracle.py compares both bit-for-bit, and tests/golden/ holds hashes generated from
 * that library by tests/golden/make_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off, no -march (reference CMake adds no FMA-enabling flags):
 * every float operation below is a separately rounded IEEE-754 binary32 operation, written in the
 * reference's evaluation order.
 *
 * Deliberate deviations (only where the reference has undefined behaviour):
 *   - solid: centre samples outside the grid are skipped (reference writes out of bounds; cannot
 *     happen when voxinfo comes from the mesh bbox);
 *   - solid: xmax < 0 flips nothing (reference CPU wraps to a ~2^32 out-of-bounds loop);
 *     xmax >= G is clamped to G-1 (reference writes out of bounds).  Both events are counted and
 *     readable through oracle_solid_ub_events() so tests can assert fixtures stay clear of them.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct { float x, y, z; } f3;
typedef struct { float x, y; } f2;

/* ------------------------------------------------------------------ helper_math.h restated */
static inline f3 f3_sub(f3 a, f3 b) { f3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }     /* :595 */
static inline f2 f2_sub(f2 a, f2 b) { f2 r = { a.x - b.x, a.y - b.y }; return r; }                 /* :526 */
static inline f2 f2_neg(f2 a) { f2 r = { -a.x, -a.y }; return r; }                                 /* :266 */
static inline float minf_(float a, float b) { return a < b ? a : b; }                              /* :58  */
static inline float maxf_(float a, float b) { return a > b ? a : b; }                              /* :63  */
static inline float dot2(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }                             /* :1260 */
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                 /* :1264 */
static inline f3 cross3(f3 a, f3 b) {                                                              /* :1436 */
	f3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
	return r;
}
static inline f3 normalize3(f3 v) {                                                                /* :1325, :78 */
	float inv_len = 1.0f / sqrtf(dot3(v, v));
	f3 r = { v.x * inv_len, v.y * inv_len, v.z * inv_len };
	return r;
}
static inline int clampi(int f, int a, int b) { int m = f < b ? f : b; return a > m ? a : m; }     /* :1172 */
/* std::max<float>(0.0f, x) as reached through timer.h's `using namespace std` */
static inline float max0(float x) { return (0.0f < x) ? x : 0.0f; }

/* ------------------------------------------------------------------ util.h restated */

/* util.h:80-110 createMeshBBCube */
void oracle_bbox_cube(const float in_min[3], const float in_max[3], float out_min[3], float out_max[3]) {
	float len[3];
	for (int k = 0; k < 3; k++) { out_min[k] = in_min[k]; out_max[k] = in_max[k]; len[k] = in_max[k] - in_min[k]; }
	float yz = len[1] > len[2] ? len[1] : len[2];       /* std::max(lengths.y, lengths.z)        */
	float max_length = len[0] < yz ? yz : len[0];       /* std::max(lengths.x, <that>)           */
	for (int k = 0; k < 3; k++) {
		if (max_length != len[k]) {
			float delta = max_length - len[k];
			out_min[k] = in_min[k] - (delta / 2.0f);
			out_max[k] = in_max[k] + (delta / 2.0f);
		}
	}
	for (int k = 0; k < 3; k++) {                        /* the 1/10001 pad, util.h:106-108       */
		float eps = (out_max[k] - out_min[k]) / 10001.0f;
		out_min[k] -= eps;
		out_max[k] += eps;
	}
}

/* min/max over all vertices (trimesh2 need_bbox, called at main.cpp:179) */
void oracle_mesh_bbox(const float* verts, size_t nv, float bb_min[3], float bb_max[3]) {
	for (int k = 0; k < 3; k++) { bb_min[k] = verts[k]; bb_max[k] = verts[k]; }
	for (size_t i = 1; i < nv; i++)
		for (int k = 0; k < 3; k++) {
			float v = verts[3 * i + k];
			if (v < bb_min[k]) bb_min[k] = v;
			if (v > bb_max[k]) bb_max[k] = v;
		}
}

/* util.h:58-60 voxinfo ctor */
void oracle_unit(const float bb_min[3], const float bb_max[3], unsigned int gridsize, float unit[3]) {
	for (int k = 0; k < 3; k++) unit[k] = (bb_max[k] - bb_min[k]) / (float)gridsize;
}

/* main.cpp:190 — the reference's size; its binary32 division rounds G^3 down for some odd grid sizes (G = 257: one bit
 * short) and the reference then writes one word past its table.  The oracle's tables are sized to hold every voxel
 * (max of both formulas) so that this out-of-bounds write of the reference lands inside the buffer and can be compared. */
size_t oracle_reference_table_bytes(unsigned int gridsize) {
	size_t g = gridsize;
	return (size_t)(ceil((g * g * g) / 32.0f) * 4);
}
size_t oracle_table_bytes(unsigned int gridsize) {
	size_t g = gridsize;
	size_t exact = ((g * g * g + 31) / 32) * 4, ref = oracle_reference_table_bytes(gridsize);
	return exact > ref ? exact : ref;
}

/* ------------------------------------------------------------------ morton (cpu_voxelizer.cpp:18-32) */

/* One entry of host_morton256_x (morton_LUTs.h:5): bit i of b goes to bit 3i. y = <<1, z = <<2. */
static inline uint32_t lut_x(uint32_t b) {
	uint32_t r = 0;
	for (int i = 0; i < 8; i++) r |= ((b >> i) & 1u) << (3 * i);
	return r;
}
uint64_t oracle_morton(unsigned int x, unsigned int y, unsigned int z) {
	uint64_t answer;
	answer = (lut_x((z >> 16) & 0xFF) << 2) | (lut_x((y >> 16) & 0xFF) << 1) | lut_x((x >> 16) & 0xFF);
	/* the reference shifts by 48 here, not 24 (cpu_voxelizer.cpp:23); irrelevant below 2^16 */
	answer = answer << 48 | (lut_x((z >> 8) & 0xFF) << 2) | (lut_x((y >> 8) & 0xFF) << 1) | lut_x((x >> 8) & 0xFF);
	answer = answer << 24 | (lut_x(z & 0xFF) << 2) | (lut_x(y & 0xFF) << 1) | lut_x(x & 0xFF);
	return answer;
}

/* ------------------------------------------------------------------ bit table (cpu_voxelizer.cpp:7-15,186-194) */
/* Test harness only: a table that holds just a z-slab of a grid too big for host memory (8192^3 = 64 GiB) starts at this word of
 * the full table; 0 (the default) = the whole table, as in the reference. */
static size_t g_table_origin = 0;
void oracle_set_table_origin(size_t first_word) { g_table_origin = first_word; }

static inline void set_bit(uint32_t* table, size_t index) {
	size_t w = index / 32 - g_table_origin;
	uint32_t mask = 1u << (31 - (uint32_t)(index % 32));
#pragma omp atomic
	table[w] |= mask;
}
static inline void xor_bit(uint32_t* table, size_t index) {
	size_t w = index / 32 - g_table_origin;
	uint32_t mask = 1u << (31 - (uint32_t)(index % 32));
#pragma omp atomic
	table[w] ^= mask;
}

static inline f3 load3(const float* p) { f3 r = { p[0], p[1], p[2] }; return r; }

/* ------------------------------------------------------------------ surface (cpu_voxelizer.cpp:35-176) */
/*
 * tris: 9 floats per triangle, model space (the layout of main.cpp:61-80).  bb_min/unit/gridsize
 * are the voxinfo fields.  z_begin/z_end restrict the z loop to [z_begin, z_end) — used to check
 * the multi-GPU slab path; pass 0, gridsize for the reference behaviour.  table is OR-ed into.
 * stats (may be NULL): [0] candidates tested, [1] setBit calls.
 */
void oracle_surface(const float* tris, size_t n_tris, const float bb_min[3], const float unit[3],
                    unsigned int gridsize, int morton, int z_begin, int z_end, uint32_t* table,
                    uint64_t* stats) {
	const f3 bbmin = { bb_min[0], bb_min[1], bb_min[2] };
	const f3 u = { unit[0], unit[1], unit[2] };
	const int gmax = (int)gridsize - 1;
	const size_t G = gridsize;
	uint64_t n_tested = 0, n_marked = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_tested, n_marked)
	for (int64_t i = 0; i < (int64_t)n_tris; i++) {
		const float* t = tris + 9 * i;
		/* :40-45 vertex shift, :63-65 gather */
		f3 v0 = f3_sub(load3(t), bbmin), v1 = f3_sub(load3(t + 3), bbmin), v2 = f3_sub(load3(t + 6), bbmin);
		f3 e0 = f3_sub(v1, v0), e1 = f3_sub(v2, v1), e2 = f3_sub(v0, v2);               /* :68-70 */
		f3 n = normalize3(cross3(e0, e1));                                             /* :72    */
		/* :76-80 grid bbox: truncating float->int of (world / unit), clamped */
		f3 lo = { minf_(v0.x, minf_(v1.x, v2.x)), minf_(v0.y, minf_(v1.y, v2.y)), minf_(v0.z, minf_(v1.z, v2.z)) };
		f3 hi = { maxf_(v0.x, maxf_(v1.x, v2.x)), maxf_(v0.y, maxf_(v1.y, v2.y)), maxf_(v0.z, maxf_(v1.z, v2.z)) };
		int x0 = clampi((int)(lo.x / u.x), 0, gmax), y0 = clampi((int)(lo.y / u.y), 0, gmax), z0 = clampi((int)(lo.z / u.z), 0, gmax);
		int x1 = clampi((int)(hi.x / u.x), 0, gmax), y1 = clampi((int)(hi.y / u.y), 0, gmax), z1 = clampi((int)(hi.z / u.z), 0, gmax);
		/* :83-87 plane test setup */
		f3 c = { 0.0f, 0.0f, 0.0f };
		if (n.x > 0.0f) c.x = u.x;
		if (n.y > 0.0f) c.y = u.y;
		if (n.z > 0.0f) c.z = u.z;
		float d1 = dot3(n, f3_sub(c, v0));
		float d2 = dot3(n, f3_sub(f3_sub(u, c), v0));
		/* :91-101 XY */
		f2 n_xy_e0 = { -1.0f * e0.y, e0.x }, n_xy_e1 = { -1.0f * e1.y, e1.x }, n_xy_e2 = { -1.0f * e2.y, e2.x };
		if (n.z < 0.0f) { n_xy_e0 = f2_neg(n_xy_e0); n_xy_e1 = f2_neg(n_xy_e1); n_xy_e2 = f2_neg(n_xy_e2); }
		f2 v0xy = { v0.x, v0.y }, v1xy = { v1.x, v1.y }, v2xy = { v2.x, v2.y };
		float d_xy_e0 = (-1.0f * dot2(n_xy_e0, v0xy)) + max0(u.x * n_xy_e0.x) + max0(u.y * n_xy_e0.y);
		float d_xy_e1 = (-1.0f * dot2(n_xy_e1, v1xy)) + max0(u.x * n_xy_e1.x) + max0(u.y * n_xy_e1.y);
		float d_xy_e2 = (-1.0f * dot2(n_xy_e2, v2xy)) + max0(u.x * n_xy_e2.x) + max0(u.y * n_xy_e2.y);
		/* :103-113 YZ */
		f2 n_yz_e0 = { -1.0f * e0.z, e0.y }, n_yz_e1 = { -1.0f * e1.z, e1.y }, n_yz_e2 = { -1.0f * e2.z, e2.y };
		if (n.x < 0.0f) { n_yz_e0 = f2_neg(n_yz_e0); n_yz_e1 = f2_neg(n_yz_e1); n_yz_e2 = f2_neg(n_yz_e2); }
		f2 v0yz = { v0.y, v0.z }, v1yz = { v1.y, v1.z }, v2yz = { v2.y, v2.z };
		float d_yz_e0 = (-1.0f * dot2(n_yz_e0, v0yz)) + max0(u.y * n_yz_e0.x) + max0(u.z * n_yz_e0.y);
		float d_yz_e1 = (-1.0f * dot2(n_yz_e1, v1yz)) + max0(u.y * n_yz_e1.x) + max0(u.z * n_yz_e1.y);
		float d_yz_e2 = (-1.0f * dot2(n_yz_e2, v2yz)) + max0(u.y * n_yz_e2.x) + max0(u.z * n_yz_e2.y);
		/* :115-125 ZX — note unit.x pairs with the p.z coefficient and unit.z with the p.x one (§A-13) */
		f2 n_zx_e0 = { -1.0f * e0.x, e0.z }, n_zx_e1 = { -1.0f * e1.x, e1.z }, n_zx_e2 = { -1.0f * e2.x, e2.z };
		if (n.y < 0.0f) { n_zx_e0 = f2_neg(n_zx_e0); n_zx_e1 = f2_neg(n_zx_e1); n_zx_e2 = f2_neg(n_zx_e2); }
		f2 v0zx = { v0.z, v0.x }, v1zx = { v1.z, v1.x }, v2zx = { v2.z, v2.x };
		float d_xz_e0 = (-1.0f * dot2(n_zx_e0, v0zx)) + max0(u.x * n_zx_e0.x) + max0(u.z * n_zx_e0.y);
		float d_xz_e1 = (-1.0f * dot2(n_zx_e1, v1zx)) + max0(u.x * n_zx_e1.x) + max0(u.z * n_zx_e1.y);
		float d_xz_e2 = (-1.0f * dot2(n_zx_e2, v2zx)) + max0(u.x * n_zx_e2.x) + max0(u.z * n_zx_e2.y);

		int za = z0 > z_begin ? z0 : z_begin, zb = z1 < z_end - 1 ? z1 : z_end - 1;
		/* :128-175 */
		for (int z = za; z <= zb; z++) {
			for (int y = y0; y <= y1; y++) {
				for (int x = x0; x <= x1; x++) {
					n_tested++;
					f3 p = { x * u.x, y * u.y, z * u.z };
					float nDOTp = dot3(n, p);
					if (((nDOTp + d1) * (nDOTp + d2)) > 0.0f) continue;
					f2 p_xy = { p.x, p.y };
					if ((dot2(n_xy_e0, p_xy) + d_xy_e0) < 0.0f) continue;
					if ((dot2(n_xy_e1, p_xy) + d_xy_e1) < 0.0f) continue;
					if ((dot2(n_xy_e2, p_xy) + d_xy_e2) < 0.0f) continue;
					f2 p_yz = { p.y, p.z };
					if ((dot2(n_yz_e0, p_yz) + d_yz_e0) < 0.0f) continue;
					if ((dot2(n_yz_e1, p_yz) + d_yz_e1) < 0.0f) continue;
					if ((dot2(n_yz_e2, p_yz) + d_yz_e2) < 0.0f) continue;
					f2 p_zx = { p.z, p.x };
					if ((dot2(n_zx_e0, p_zx) + d_xz_e0) < 0.0f) continue;
					if ((dot2(n_zx_e1, p_zx) + d_xz_e1) < 0.0f) continue;
					if ((dot2(n_zx_e2, p_zx) + d_xz_e2) < 0.0f) continue;
					n_marked++;
					size_t location = morton ? (size_t)oracle_morton(x, y, z)
					                         : (size_t)x + (size_t)y * G + (size_t)z * G * G;   /* :164,168 */
					set_bit(table, location);
				}
			}
		}
	}
	if (stats) { stats[0] = n_tested; stats[1] = n_marked; }
}

/* ------------------------------------------------------------------ solid helpers (cpu_voxelizer.cpp:196-238) */
static inline int top_left_edge(f2 v0, f2 v1) {                                                    /* :196 */
	return ((v1.y < v0.y) || (v1.y == v0.y && v0.x > v1.x));
}
static inline int check_ccw(f2 v0, f2 v1, f2 v2) {                                                 /* :201 */
	f2 e0 = f2_sub(v1, v0), e1 = f2_sub(v2, v0);
	float result = e0.x * e1.y - e1.x * e0.y;
	return result > 0;
}
static inline float get_x_coordinate(f3 n, f3 v0, f2 point) {                                      /* :212 */
	return (-(n.y * (point.x - v0.y) + n.z * (point.y - v0.z)) / n.x + v0.x);
}
#define FLOAT_ERROR 0.000001   /* double literal, :2 — comparisons against it promote to double */
static inline int check_point_triangle(f2 v0, f2 v1, f2 v2, f2 point) {                            /* :217 */
	f2 PA = f2_sub(point, v0), PB = f2_sub(point, v1), PC = f2_sub(point, v2);
	float t1 = PA.x * PB.y - PA.y * PB.x;
	if ((double)fabsf(t1) < FLOAT_ERROR && PA.x * PB.x <= 0 && PA.y * PB.y <= 0) return 1;
	float t2 = PB.x * PC.y - PB.y * PC.x;
	if ((double)fabsf(t2) < FLOAT_ERROR && PB.x * PC.x <= 0 && PB.y * PC.y <= 0) return 2;
	float t3 = PC.x * PA.y - PC.y * PA.x;
	if ((double)fabsf(t3) < FLOAT_ERROR && PC.x * PA.x <= 0 && PC.y * PA.y <= 0) return 3;
	if (t1 * t2 > 0 && t1 * t3 > 0) return 0;
	return -1;
}

static uint64_t g_solid_ub_events = 0;
uint64_t oracle_solid_ub_events(void) { return g_solid_ub_events; }

/* ------------------------------------------------------------------ solid (cpu_voxelizer.cpp:241-312) */
/*
 * z_begin/z_end: restrict centre-sample z to [z_begin, z_end) (multi-GPU slab check).
 * stats (may be NULL): [0] column hits (accepted (y,z) samples), [1] bit flips.
 */
void oracle_solid(const float* tris, size_t n_tris, const float bb_min[3], const float unit[3],
                  unsigned int gridsize, int morton, int z_begin, int z_end, uint32_t* table,
                  uint64_t* stats) {
	const f3 bbmin = { bb_min[0], bb_min[1], bb_min[2] };
	const f3 u = { unit[0], unit[1], unit[2] };
	const size_t G = gridsize;
	uint64_t n_hits = 0, n_flips = 0, n_ub = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_hits, n_flips, n_ub)
	for (int64_t i = 0; i < (int64_t)n_tris; i++) {
		const float* t = tris + 9 * i;
		f3 v0 = f3_sub(load3(t), bbmin), v1 = f3_sub(load3(t + 3), bbmin), v2 = f3_sub(load3(t + 6), bbmin);
		f3 e0 = f3_sub(v1, v0), e1 = f3_sub(v2, v1);                                    /* :260-261 */
		f3 n = normalize3(cross3(e0, e1));                                             /* :264 */
		if ((double)fabsf(n.x) < FLOAT_ERROR) continue;                                /* :265 */
		f2 v0_yz = { v0.y, v0.z }, v1_yz = { v1.y, v1.z }, v2_yz = { v2.y, v2.z };     /* :268-270 */
		if (!check_ccw(v0_yz, v1_yz, v2_yz)) { f2 v3 = v1_yz; v1_yz = v2_yz; v2_yz = v3; }  /* :273-278 */
		f2 bbox_max = { maxf_(v0_yz.x, maxf_(v1_yz.x, v2_yz.x)), maxf_(v0_yz.y, maxf_(v1_yz.y, v2_yz.y)) };
		f2 bbox_min = { minf_(v0_yz.x, minf_(v1_yz.x, v2_yz.x)), minf_(v0_yz.y, minf_(v1_yz.y, v2_yz.y)) };
		/* :285-286 centre-sample bbox, all binary32 */
		f2 bbox_max_grid = { floorf(bbox_max.x / u.y - 0.5f), floorf(bbox_max.y / u.z - 0.5f) };
		f2 bbox_min_grid = { ceilf(bbox_min.x / u.y - 0.5f), ceilf(bbox_min.y / u.z - 0.5f) };
		for (int y = (int)bbox_min_grid.x; (float)y <= bbox_max_grid.x; y++) {          /* :288 */
			for (int z = (int)bbox_min_grid.y; (float)z <= bbox_max_grid.y; z++) {      /* :290 */
				if (z < z_begin || z >= z_end) continue;
				if (y < 0 || y >= (int)G || z < 0 || z >= (int)G) { n_ub++; continue; }  /* only if voxinfo does not enclose the mesh */
				f2 point = { (y + 0.5f) * u.y, (z + 0.5f) * u.z };                       /* :292 */
				int checknum = check_point_triangle(v0_yz, v1_yz, v2_yz, point);
				if ((checknum == 1 && top_left_edge(v0_yz, v1_yz)) || (checknum == 2 && top_left_edge(v1_yz, v2_yz)) ||
				    (checknum == 3 && top_left_edge(v2_yz, v0_yz)) || (checknum == 0)) {
					/* :296 — float / float, then double subtraction of 0.5, then int() truncation */
					int xmax = (int)(get_x_coordinate(n, v0, point) / u.x - 0.5);
					n_hits++;
					if (xmax < 0) { n_ub++; continue; }   /* value <= -1: reference CPU UB; (-1,0] already truncated to 0 */
					if (xmax > (int)G - 1) { xmax = (int)G - 1; n_ub++; }
					for (int x = 0; x <= xmax; x++) {                                    /* :297-308 */
						size_t location = morton ? (size_t)oracle_morton(x, y, z)
						                         : (size_t)x + (size_t)y * G + (size_t)z * G * G;
						xor_bit(table, location);
					}
					n_flips += (uint64_t)xmax + 1;
				}
			}
		}
	}
	if (stats) { stats[0] = n_hits; stats[1] = n_flips; }
#pragma omp atomic
	g_solid_ub_events += n_ub;
}

/* ------------------------------------------------------------------ conveniences for the checker */

/* Expand an indexed mesh to the 9-float soup of main.cpp:61-80. */
void oracle_expand_soup(const float* verts, const int32_t* faces, size_t nf, float* tris) {
	for (size_t i = 0; i < nf; i++)
		for (int k = 0; k < 3; k++) memcpy(tris + 9 * i + 3 * k, verts + 3 * (size_t)faces[3 * i + k], 3 * sizeof(float));
}

/* FNV-1a-64 over the table bytes in memory order (the hash SURVEY.md §8c quotes). */
uint64_t oracle_fnv1a64(const void* data, size_t n) {
	const unsigned char* p = (const unsigned char*)data;
	uint64_t h = 0xcbf29ce484222325ULL;
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
	return h;
}

uint64_t oracle_popcount(const uint32_t* table, size_t n_words) {
	uint64_t c = 0;
	for (size_t i = 0; i < n_words; i++) c += (uint64_t)__builtin_popcount(table[i]);
	return c;
}

/* util.h:25-38 checkVoxel — the reader side of the layout contract */
int oracle_check_voxel(size_t x, size_t y, size_t z, unsigned int gridsize, const uint32_t* table) {
	size_t location = x + (y * gridsize) + (z * (size_t)gridsize * gridsize);
	size_t w = location / 32;
	unsigned int bit_pos = 31 - (unsigned int)(location % 32);
	return (table[w] & (1u << bit_pos)) != 0;
}
