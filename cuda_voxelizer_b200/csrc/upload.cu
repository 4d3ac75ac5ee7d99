// upload.cu — device side of the triangle upload path (replaces the serial host gather of
// meshToGPU_managed, main.cpp:61-80): index expansion, AoS -> SoA float4 transposition and the
// mesh bbox reduction (trimesh2 need_bbox, main.cpp:179) run on the GPU, so the PCIe transfer is
// the indexed mesh (12 B/vertex + 12 B/face) instead of the 36 B/triangle soup.
#include "vox_internal.h"

namespace voxb {

constexpr int kUpBlock = 256;

__global__ void __launch_bounds__(kUpBlock) soup_to_soa4_kernel(const float* __restrict__ soup, float4* __restrict__ out, size_t n) {
	const size_t i = (size_t)blockIdx.x * kUpBlock + threadIdx.x;
	if (i >= n) return;
	const float* p = soup + 9 * i;
	out[i] = make_float4(__ldg(p + 0), __ldg(p + 1), __ldg(p + 2), 0.0f);
	out[n + i] = make_float4(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5), 0.0f);
	out[2 * n + i] = make_float4(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8), 0.0f);
}

template <bool SOA4>
// Face indices outside [0, n_verts) are counted into *bad (the caller reports them) and clamped, so nothing is read out of bounds.
__global__ void __launch_bounds__(kUpBlock) expand_indexed_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                                                                  size_t n_faces, unsigned int n_verts, unsigned long long* __restrict__ bad,
                                                                  float* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * kUpBlock + threadIdx.x;
	if (i >= n_faces) return;
	int a = __ldg(faces + 3 * i), b = __ldg(faces + 3 * i + 1), c = __ldg(faces + 3 * i + 2);
	if (n_verts) {
		const bool oob = (unsigned int)a >= n_verts || (unsigned int)b >= n_verts || (unsigned int)c >= n_verts;
		if (oob) {
			if (bad) atomicAdd(bad, 1ull);
			a = min(max(a, 0), (int)n_verts - 1); b = min(max(b, 0), (int)n_verts - 1); c = min(max(c, 0), (int)n_verts - 1);
		}
	}
	const float* pa = verts + 3 * (size_t)a;
	const float* pb = verts + 3 * (size_t)b;
	const float* pc = verts + 3 * (size_t)c;
	if (SOA4) {
		float4* o = reinterpret_cast<float4*>(out);
		o[i] = make_float4(__ldg(pa), __ldg(pa + 1), __ldg(pa + 2), 0.0f);
		o[n_faces + i] = make_float4(__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), 0.0f);
		o[2 * n_faces + i] = make_float4(__ldg(pc), __ldg(pc + 1), __ldg(pc + 2), 0.0f);
	} else {
		float* o = out + 9 * i;
		o[0] = __ldg(pa); o[1] = __ldg(pa + 1); o[2] = __ldg(pa + 2);
		o[3] = __ldg(pb); o[4] = __ldg(pb + 1); o[5] = __ldg(pb + 2);
		o[6] = __ldg(pc); o[7] = __ldg(pc + 1); o[8] = __ldg(pc + 2);
	}
}

// order-preserving float <-> uint key, so min/max can use integer atomics
__device__ __forceinline__ unsigned int fkey(float f) {
	const unsigned int b = __float_as_uint(f);
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funkey(unsigned int k) {
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void bbox_init_kernel(unsigned int* keys) {
	if (threadIdx.x < 3) keys[threadIdx.x] = 0xffffffffu;
	else if (threadIdx.x < 6) keys[threadIdx.x] = 0u;
}
__global__ void __launch_bounds__(kUpBlock) bbox_reduce_kernel(const float* __restrict__ verts, size_t n_verts, unsigned int* keys) {
	unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
	const size_t stride = (size_t)gridDim.x * kUpBlock;
	for (size_t i = (size_t)blockIdx.x * kUpBlock + threadIdx.x; i < n_verts; i += stride) {
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const unsigned int key = fkey(__ldg(verts + 3 * i + k));
			lo[k] = min(lo[k], key);
			hi[k] = max(hi[k], key);
		}
	}
#pragma unroll
	for (int k = 0; k < 3; k++) {
		lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
		hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int k = 0; k < 3; k++) { atomicMin(keys + k, lo[k]); atomicMax(keys + 3 + k, hi[k]); }
	}
}
__global__ void bbox_decode_kernel(unsigned int* keys) {
	if (threadIdx.x < 6) {
		const float f = funkey(keys[threadIdx.x]);
		reinterpret_cast<float*>(keys)[threadIdx.x] = f;
	}
}

// Triangle routing for multi-GPU regions: keeps the triangles whose footprint can touch `g`'s
// region — surface: the clamped grid bbox of voxelize.cu:86-87; solid: the centre-sample (y,z) range of
// voxelize_solid.cu:112-113 — using the same arithmetic as the voxelization kernels, and appends them to a
// compact soup through a warp-aggregated cursor.
template <bool SOLID>
__global__ void __launch_bounds__(kUpBlock) route_kernel(const GridParams g, const float* __restrict__ soup, float* __restrict__ out,
                                                         unsigned long long* __restrict__ cursor) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kUpBlock + threadIdx.x;
	bool keep = false;
	Tri t;
	if (i < g.n_tris) {
		load_tri_aos(soup, i, t);
		Tri ts = t;
		shift_tri(ts, g);
		if (SOLID) {
			SolidSetup s;
			solid_setup(ts, g, s);
			keep = !s.skip && max(s.y0, g.ry0) <= min(s.y1, g.ry1 - 1) && max(s.z0, g.rz0) <= min(s.z1, g.rz1 - 1);
		} else {
			SurfSetup s;
			surf_bbox(ts, g, s);
			keep = max(s.x0, g.rx0) <= min(s.x1, g.rx1 - 1) && max(s.y0, g.ry0) <= min(s.y1, g.ry1 - 1) && max(s.z0, g.rz0) <= min(s.z1, g.rz1 - 1);
		}
	}
	const unsigned int m = __ballot_sync(0xffffffffu, keep);
	if (m == 0u) return;
	const int lane = threadIdx.x & 31;
	unsigned long long base = 0ull;
	if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (keep) {
		float* o = out + 9ull * (base + __popc(m & ((1u << lane) - 1u)));
		o[0] = t.v0x; o[1] = t.v0y; o[2] = t.v0z; o[3] = t.v1x; o[4] = t.v1y; o[5] = t.v1z; o[6] = t.v2x; o[7] = t.v2y; o[8] = t.v2z;
	}
}

// Multi-region routing (N ranks): pass 1 computes, per triangle, the bit mask of the regions it can touch and counts
// per region; pass 2 appends each triangle to the segment of every region in its mask.  Regions are boxes.
struct RouteBoxes {
	int lo[32][3];
	int hi[32][3];
	int n;
};

template <bool SOLID>
__device__ __forceinline__ unsigned int route_mask(const GridParams& g, const RouteBoxes& rb, const Tri& t_model) {
	Tri ts = t_model;
	shift_tri(ts, g);
	int x0, x1, y0, y1, z0, z1;
	if (SOLID) {
		SolidSetup s;
		solid_setup(ts, g, s);
		if (s.skip) return 0u;
		x0 = 0; x1 = g.G - 1; y0 = s.y0; y1 = s.y1; z0 = s.z0; z1 = s.z1;
	} else {
		SurfSetup s;
		surf_bbox(ts, g, s);
		x0 = s.x0; x1 = s.x1; y0 = s.y0; y1 = s.y1; z0 = s.z0; z1 = s.z1;
	}
	unsigned int m = 0u;
	for (int r = 0; r < rb.n; r++) {
		const bool hit = max(x0, rb.lo[r][0]) <= min(x1, rb.hi[r][0] - 1) && max(y0, rb.lo[r][1]) <= min(y1, rb.hi[r][1] - 1) &&
		                 max(z0, rb.lo[r][2]) <= min(z1, rb.hi[r][2] - 1);
		m |= hit ? (1u << r) : 0u;
	}
	return m;
}

template <bool SOLID>
__global__ void __launch_bounds__(kUpBlock) route_count_kernel(const GridParams g, const RouteBoxes rb, const float* __restrict__ soup,
                                                               unsigned int* __restrict__ masks, unsigned long long* __restrict__ counts) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kUpBlock + threadIdx.x;
	unsigned int m = 0u;
	if (i < g.n_tris) {
		Tri t;
		load_tri_aos(soup, i, t);
		m = route_mask<SOLID>(g, rb, t);
		masks[i] = m;
	}
	for (int r = 0; r < rb.n; r++) {
		const unsigned int b = __ballot_sync(0xffffffffu, (m >> r) & 1u);
		if (b && (threadIdx.x & 31) == 0) atomicAdd(counts + r, (unsigned long long)__popc(b));
	}
}

__global__ void __launch_bounds__(kUpBlock) route_scatter_kernel(unsigned long long n_tris, int n_regions, const float* __restrict__ soup,
                                                                 const unsigned int* __restrict__ masks, float* __restrict__ out,
                                                                 unsigned long long* __restrict__ cursors /* pre-set to segment starts */) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kUpBlock + threadIdx.x;
	const unsigned int m = i < n_tris ? masks[i] : 0u;
	const int lane = threadIdx.x & 31;
	float v[9];
	if (m) {
#pragma unroll
		for (int k = 0; k < 9; k++) v[k] = __ldg(soup + 9ull * i + k);
	}
	for (int r = 0; r < n_regions; r++) {
		const bool mine = (m >> r) & 1u;
		const unsigned int b = __ballot_sync(0xffffffffu, mine);
		if (!b) continue;
		unsigned long long base = 0ull;
		if (lane == 0) base = atomicAdd(cursors + r, (unsigned long long)__popc(b));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (mine) {
			float* o = out + 9ull * (base + __popc(b & ((1u << lane) - 1u)));
#pragma unroll
			for (int k = 0; k < 9; k++) o[k] = v[k];
		}
	}
}

cudaError_t launch_route_count(const GridParams& g, bool solid, const int (*lo)[3], const int (*hi)[3], int n_regions, const float* d_soup,
                               unsigned int* d_masks, unsigned long long* d_counts, cudaStream_t st) {
	RouteBoxes rb;
	rb.n = n_regions;
	for (int r = 0; r < n_regions; r++) for (int k = 0; k < 3; k++) { rb.lo[r][k] = lo[r][k]; rb.hi[r][k] = hi[r][k]; }
	cudaError_t err = cudaMemsetAsync(d_counts, 0, 32 * sizeof(unsigned long long), st);
	if (err != cudaSuccess || g.n_tris == 0) return err;
	const unsigned blocks = (unsigned)((g.n_tris + kUpBlock - 1) / kUpBlock);
	if (solid) route_count_kernel<true><<<blocks, kUpBlock, 0, st>>>(g, rb, d_soup, d_masks, d_counts);
	else route_count_kernel<false><<<blocks, kUpBlock, 0, st>>>(g, rb, d_soup, d_masks, d_counts);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_route_scatter(unsigned long long n_tris, int n_regions, const float* d_soup, const unsigned int* d_masks, float* d_out,
                                 unsigned long long* d_cursors, cudaStream_t st) {
	if (n_tris == 0) return cudaSuccess;
	const unsigned blocks = (unsigned)((n_tris + kUpBlock - 1) / kUpBlock);
	route_scatter_kernel<<<blocks, kUpBlock, 0, st>>>(n_tris, n_regions, d_soup, d_masks, d_out, d_cursors);
	g_launch_count++;
	return cudaGetLastError();
}

// Layer sort (upload path).  A soup that will be voxelized more than once is worth ordering by the z-layer of each
// triangle's lowest vertex: the per-triangle kernel's atomics then sweep the table front to back (z is the slowest
// index of the linear order), the sectors they touch are fetched once and stay in L2, and the DRAM read-modify-write
// that otherwise costs a third of the kernel disappears (10M triangles @2048^3: 0.496 -> 0.356 ms, same table bits:
// OR / XOR do not depend on the triangle order).  Counting sort: key + histogram, scan, scatter; the order inside a
// layer is whatever the atomics make it.
__global__ void __launch_bounds__(kUpBlock) layer_key_kernel(const GridParams g, const float* __restrict__ soup, unsigned int* __restrict__ keys,
                                                             unsigned int* __restrict__ hist) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kUpBlock + threadIdx.x;
	if (i >= g.n_tris) return;
	const float* p = soup + 9ull * i;
	const float zmin = fminf(__ldg(p + 2), fminf(__ldg(p + 5), __ldg(p + 8)));
	const float q = (zmin - g.bz) * g.ruz;                         // ordering only: any monotone key will do
	const unsigned int key = (unsigned int)max(0, min(g.G - 1, q > 0.0f ? __float2int_rd(q) : 0));
	keys[i] = key;
	const unsigned int peers = __match_any_sync(__activemask(), key);
	if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + key, (unsigned int)__popc(peers));
}
__global__ void layer_scan_kernel(unsigned int* __restrict__ hist, int n) {          // counts -> exclusive offsets, in place
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	unsigned int run = 0u;
	for (int k = 0; k < n; k++) { const unsigned int c = hist[k]; hist[k] = run; run += c; }
}
__global__ void __launch_bounds__(kUpBlock) layer_scatter_kernel(unsigned long long n_tris, const float* __restrict__ soup, const unsigned int* __restrict__ keys,
                                                                 unsigned int* __restrict__ cursor, float* __restrict__ out) {
	const unsigned long long i = (unsigned long long)blockIdx.x * kUpBlock + threadIdx.x;
	if (i >= n_tris) return;
	const unsigned int key = keys[i];
	const unsigned int peers = __match_any_sync(__activemask(), key);
	const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
	unsigned int base = 0u;
	if (lane == leader) base = atomicAdd(cursor + key, (unsigned int)__popc(peers));
	base = __shfl_sync(peers, base, leader);
	const float* p = soup + 9ull * i;
	float* o = out + 9ull * ((unsigned long long)base + __popc(peers & ((1u << lane) - 1u)));
#pragma unroll
	for (int k = 0; k < 9; k++) o[k] = __ldg(p + k);
}

cudaError_t launch_layer_sort(const GridParams& g, const float* d_soup, float* d_out, unsigned int* d_keys, unsigned int* d_hist, cudaStream_t st) {
	cudaError_t e = cudaMemsetAsync(d_hist, 0, (size_t)g.G * sizeof(unsigned int), st);
	if (e != cudaSuccess || g.n_tris == 0) return e;
	const unsigned int blocks = (unsigned int)((g.n_tris + kUpBlock - 1) / kUpBlock);
	layer_key_kernel<<<blocks, kUpBlock, 0, st>>>(g, d_soup, d_keys, d_hist);
	layer_scan_kernel<<<1, 32, 0, st>>>(d_hist, g.G);
	layer_scatter_kernel<<<blocks, kUpBlock, 0, st>>>(g.n_tris, d_soup, d_keys, d_hist, d_out);
	g_launch_count += 3;
	return cudaGetLastError();
}

cudaError_t launch_route(const GridParams& g, bool solid, const float* d_soup, float* d_out, unsigned long long* d_cursor, cudaStream_t st) {
	cudaError_t err = cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st);
	if (err != cudaSuccess || g.n_tris == 0) return err;
	const unsigned blocks = (unsigned)((g.n_tris + kUpBlock - 1) / kUpBlock);
	if (solid) route_kernel<true><<<blocks, kUpBlock, 0, st>>>(g, d_soup, d_out, d_cursor);
	else route_kernel<false><<<blocks, kUpBlock, 0, st>>>(g, d_soup, d_out, d_cursor);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_soup_to_soa4(const float* d_soup, float* d_soa4, size_t n_tris, cudaStream_t st) {
	if (n_tris == 0) return cudaSuccess;
	soup_to_soa4_kernel<<<(unsigned)((n_tris + kUpBlock - 1) / kUpBlock), kUpBlock, 0, st>>>(d_soup, reinterpret_cast<float4*>(d_soa4), n_tris);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_expand_indexed(const float* d_verts, const int* d_faces, size_t n_faces, size_t n_verts,
                                  bool soa4, float* d_out, cudaStream_t st, unsigned long long* d_bad) {
	if (n_faces == 0) return cudaSuccess;
	const unsigned blocks = (unsigned)((n_faces + kUpBlock - 1) / kUpBlock);
	const unsigned int nv = n_verts > 0x7fffffffull ? 0x7fffffffu : (unsigned int)n_verts;       // 0: the caller vouches for the indices
	if (soa4) expand_indexed_kernel<true><<<blocks, kUpBlock, 0, st>>>(d_verts, d_faces, n_faces, nv, d_bad, d_out);
	else expand_indexed_kernel<false><<<blocks, kUpBlock, 0, st>>>(d_verts, d_faces, n_faces, nv, d_bad, d_out);
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_bbox_reduce(const float* d_verts, size_t n_verts, float* d_minmax6, cudaStream_t st) {
	unsigned int* keys = reinterpret_cast<unsigned int*>(d_minmax6);
	bbox_init_kernel<<<1, 32, 0, st>>>(keys);
	size_t blocks = (n_verts + kUpBlock - 1) / kUpBlock;
	if (blocks > 148 * 8) blocks = 148 * 8;
	bbox_reduce_kernel<<<(unsigned)blocks, kUpBlock, 0, st>>>(d_verts, n_verts, keys);
	bbox_decode_kernel<<<1, 32, 0, st>>>(keys);
	g_launch_count += 3;
	return cudaGetLastError();
}

}  // namespace voxb
