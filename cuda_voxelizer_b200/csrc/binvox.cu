// binvox.cu — the binvox payload on the device (replaces the G^3 checkVoxel loop of util_io.cpp:219-245).
//
// The format visits the voxels x-major, then z, then y and run-length encodes them as (value, count <= 255) byte pairs.  The
// table is x-fastest, so:
//   1. binvox_edges_kernel    the table re-ordered to the traversal — stream position k = (x G + z) G + y — and differentiated in
//                             one pass: bit k of the EDGE table = (voxel k != voxel k-1), bit 0 always set = "a run starts here".
//                             A block moves a 256(x) x 256(y) patch of one z-layer: whole 32-byte sectors in (8 x-words per row)
//                             and out (8 y-words per x); 32x32 bit blocks are turned by a five-stage shuffle butterfly.
//   2. extract_* (extract.cu) the edge table's set bits as ascending positions: run i starts at P[i].
//   3. binvox_count_kernel    pairs per run, ceil(length / 255), summed per block; binvox_scan_kernel (one CTA) -> block offsets.
//   4. binvox_emit_kernel     every run writes its pairs: (v, 255) ... (v, remainder) — exactly what the reference's counter
//                             produces (it flushes when the count reaches 255 or the value changes).  Runs alternate in value:
//                             run i has value v0 ^ (i & 1).  Short runs are written by their lane, longer ones by the warp,
//                             runs of more than 2^16 pairs (large empty regions) by binvox_long_kernel, grid-wide.
// Output: 2 bytes per pair in a device buffer; the caller copies it out behind the ASCII header (util_io.cpp:210-216).
#include "../../include/voxb200.h"
#include "vox_internal.h"

namespace voxb {

namespace {

constexpr int kBvBlock = 256;
constexpr int kRunsPerThread = 4;
constexpr int kRunsPerBlock = kBvBlock * kRunsPerThread;
constexpr unsigned int kLongRunPairs = 1u << 16;

#define BV_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); if (d_out) cudaFree(d_out); return abi_fail_cuda(e_, #call); } } while (0)

// 32x32 bit transpose across a warp: on entry lane r holds row r (bit c = element (r, c), c counted from the LSB), on exit lane r
// holds column r (bit c = element (c, r)).  Five exchange stages, block size 16, 8, 4, 2, 1.
__device__ __forceinline__ unsigned int transpose32(unsigned int x, int lane) {
	const unsigned int masks[5] = {0xffff0000u, 0xff00ff00u, 0xf0f0f0f0u, 0xccccccccu, 0xaaaaaaaau};
#pragma unroll
	for (int s = 0; s < 5; s++) {
		const int j = 16 >> s;
		const unsigned int m = masks[s];
		const unsigned int y = __shfl_xor_sync(0xffffffffu, x, j);
		x = (lane & j) ? ((x & m) | ((y >> j) & ~m)) : ((x & ~m) | ((y << j) & m));
	}
	return x;
}

__device__ __forceinline__ unsigned int voxel_bit(const unsigned int* __restrict__ table, unsigned int G, unsigned int x, unsigned int y, unsigned int z) {
	const size_t w = (size_t)(x >> 5) + (size_t)(G >> 5) * ((size_t)y + (size_t)G * z);
	return (__ldg(table + w) >> (31u - (x & 31u))) & 1u;
}

// grid (G/256, G/256, G): patch (tx, ty) of layer z.  `edges` holds G^3 bits in stream order, MSB first.
__global__ void __launch_bounds__(kBvBlock) binvox_edges_kernel(const unsigned int* __restrict__ table, unsigned int G, unsigned int* __restrict__ edges) {
	__shared__ unsigned int patch[256 * 9];               // 256 rows x 8 words, padded to 9
	const unsigned int tx = blockIdx.x, ty = blockIdx.y, z = blockIdx.z;
	const unsigned int Gw = G >> 5;
	{
		const unsigned int y = ty * 256u + threadIdx.x;
		const uint4* src = reinterpret_cast<const uint4*>(table + ((size_t)z * G + y) * Gw + tx * 8u);
		const uint4 a = __ldg(src), b = __ldg(src + 1);
		unsigned int* row = patch + threadIdx.x * 9;
		row[0] = a.x; row[1] = a.y; row[2] = a.z; row[3] = a.w; row[4] = b.x; row[5] = b.y; row[6] = b.z; row[7] = b.w;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;      // warp w turns x-word w of the patch
	unsigned int out[8];
#pragma unroll
	for (int yw = 0; yw < 8; yw++) {
		// lane = y within the group, bit (31 - xi) = x; after the turn lane r holds x = 31 - r with bit c = y c: reverse to MSB-first y
		out[yw] = __brev(transpose32(patch[(yw * 32 + lane) * 9 + w], lane));
	}
	const unsigned int x = tx * 256u + (unsigned int)w * 32u + (31u - (unsigned int)lane);
	// the voxel in front of this lane's first one in stream order
	unsigned int prev;
	if (ty > 0) prev = voxel_bit(table, G, x, ty * 256u - 1u, z);
	else if (z > 0) prev = voxel_bit(table, G, x, G - 1u, z - 1u);
	else if (x > 0) prev = voxel_bit(table, G, x - 1u, G - 1u, G - 1u);
	else prev = (~out[0]) >> 31;                                  // k = 0: always a run start
	unsigned int e[8];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		e[j] = out[j] ^ ((out[j] >> 1) | (prev << 31));
		prev = out[j] & 1u;
	}
	uint4* dst = reinterpret_cast<uint4*>(edges + ((size_t)x * G + z) * Gw + ty * 8u);
	dst[0] = make_uint4(e[0], e[1], e[2], e[3]);
	dst[1] = make_uint4(e[4], e[5], e[6], e[7]);
}

__device__ __forceinline__ unsigned long long run_pairs(const unsigned long long* __restrict__ starts, unsigned long long n_runs, unsigned long long total,
                                                        unsigned long long i, unsigned long long& len) {
	const unsigned long long a = starts[i], b = i + 1 < n_runs ? starts[i + 1] : total;
	len = b - a;
	return (len + 254ull) / 255ull;
}

__device__ __forceinline__ unsigned long long block_scan_u64(unsigned long long v, unsigned long long* smem, unsigned long long& total) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned long long inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += up;
	}
	if (lane == 31) smem[wid] = inc;
	__syncthreads();
	unsigned long long base = 0;
	total = 0;
#pragma unroll
	for (int k = 0; k < kBvBlock / 32; k++) {
		const unsigned long long s = smem[k];
		if (k < wid) base += s;
		total += s;
	}
	__syncthreads();
	return base + inc - v;          // exclusive
}

__global__ void __launch_bounds__(kBvBlock) binvox_count_kernel(const unsigned long long* __restrict__ starts, unsigned long long n_runs, unsigned long long total,
                                                                unsigned long long* __restrict__ block_sums) {
	__shared__ unsigned long long smem[kBvBlock / 32];
	const unsigned long long first = ((unsigned long long)blockIdx.x * kBvBlock + threadIdx.x) * kRunsPerThread;
	unsigned long long c = 0, len;
#pragma unroll
	for (int k = 0; k < kRunsPerThread; k++) if (first + k < n_runs) c += run_pairs(starts, n_runs, total, first + k, len);
	unsigned long long sum;
	block_scan_u64(c, smem, sum);
	if (threadIdx.x == 0) block_sums[blockIdx.x] = sum;
}

// one CTA: exclusive scan of n block sums in place; sums[n] = total
__global__ void __launch_bounds__(1024) binvox_scan_kernel(unsigned long long* __restrict__ sums, size_t n) {
	__shared__ unsigned long long part[1024];
	const size_t per = (n + 1023) / 1024;
	const size_t a = (size_t)threadIdx.x * per < n ? (size_t)threadIdx.x * per : n, b = a + per < n ? a + per : n;
	unsigned long long s = 0;
	for (size_t i = a; i < b; i++) s += sums[i];
	part[threadIdx.x] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long run = 0;
		for (int t = 0; t < 1024; t++) { const unsigned long long v = part[t]; part[t] = run; run += v; }
		sums[n] = run;
	}
	__syncthreads();
	unsigned long long run = part[threadIdx.x];
	for (size_t i = a; i < b; i++) { const unsigned long long v = sums[i]; sums[i] = run; run += v; }
}

struct LongRun { unsigned long long first_pair, pairs, len; unsigned int value, pad; };

__device__ __forceinline__ unsigned short pair_of(unsigned int value, unsigned long long q, unsigned long long pairs, unsigned long long len) {
	const unsigned int count = q + 1 < pairs ? 255u : (unsigned int)(len - 255ull * (pairs - 1));
	return (unsigned short)(value | (count << 8));           // byte 0 = value, byte 1 = count
}

__global__ void __launch_bounds__(kBvBlock) binvox_emit_kernel(const unsigned long long* __restrict__ starts, unsigned long long n_runs, unsigned long long total,
                                                               const unsigned long long* __restrict__ block_offsets, unsigned int v0,
                                                               unsigned short* __restrict__ out, LongRun* __restrict__ long_runs, unsigned int* __restrict__ n_long) {
	__shared__ unsigned long long smem[kBvBlock / 32];
	const unsigned long long first = ((unsigned long long)blockIdx.x * kBvBlock + threadIdx.x) * kRunsPerThread;
	unsigned long long c[kRunsPerThread], len[kRunsPerThread], sum = 0;
#pragma unroll
	for (int k = 0; k < kRunsPerThread; k++) {
		c[k] = 0; len[k] = 0;
		if (first + k < n_runs) c[k] = run_pairs(starts, n_runs, total, first + k, len[k]);
		sum += c[k];
	}
	unsigned long long block_total;
	unsigned long long at = block_offsets[blockIdx.x] + block_scan_u64(sum, smem, block_total);
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int k = 0; k < kRunsPerThread; k++) {
		const unsigned int value = (v0 ^ (unsigned int)((first + k) & 1ull)) & 1u;
		const unsigned long long pairs = c[k];
		if (pairs > kLongRunPairs) {
			LongRun r;
			r.first_pair = at; r.pairs = pairs; r.len = len[k]; r.value = value; r.pad = 0u;
			long_runs[atomicAdd(n_long, 1u)] = r;
		} else if (pairs <= 4) {
			for (unsigned long long q = 0; q < pairs; q++) out[at + q] = pair_of(value, q, pairs, len[k]);
		}
		// runs of 5 .. 2^16 pairs: the whole warp writes each of them
		unsigned int todo = __ballot_sync(0xffffffffu, pairs > 4 && pairs <= kLongRunPairs);
		while (todo) {
			const int src = __ffs(todo) - 1;
			todo &= todo - 1u;
			const unsigned long long r_at = __shfl_sync(0xffffffffu, at, src), r_pairs = __shfl_sync(0xffffffffu, pairs, src), r_len = __shfl_sync(0xffffffffu, len[k], src);
			const unsigned int r_value = __shfl_sync(0xffffffffu, value, src);
			for (unsigned long long q = lane; q < r_pairs; q += 32) out[r_at + q] = pair_of(r_value, q, r_pairs, r_len);
		}
		at += pairs;
	}
}

__global__ void __launch_bounds__(kBvBlock) binvox_long_kernel(const LongRun* __restrict__ long_runs, const unsigned int* __restrict__ n_long, unsigned short* __restrict__ out) {
	const unsigned int n = *n_long;
	const unsigned long long tid = (unsigned long long)blockIdx.x * kBvBlock + threadIdx.x, stride = (unsigned long long)gridDim.x * kBvBlock;
	for (unsigned int r = 0; r < n; r++) {
		const LongRun R = long_runs[r];
		for (unsigned long long q = tid; q < R.pairs; q += stride) out[R.first_pair + q] = pair_of(R.value, q, R.pairs, R.len);
	}
}

}  // namespace

}  // namespace voxb

using namespace voxb;

extern "C" int voxb200_binvox_rle(const unsigned int* d_table, unsigned int gridsize, unsigned char** d_bytes, size_t* n_bytes, void* stream) {
	if (!d_table || !d_bytes || !n_bytes) return abi_fail(VOXB200_EINVAL, "NULL pointer");
	const unsigned int G = gridsize;
	if (G < 256 || (G % 256u) != 0 || G > 4096) return abi_fail(VOXB200_EINVAL, "voxb200_binvox_rle needs a grid size that is a multiple of 256, at most 4096 (got %u)", G);
	if (reinterpret_cast<uintptr_t>(d_table) & 15u) return abi_fail(VOXB200_EINVAL, "the table must be 16-byte aligned");
	Workspace* ws;
	int rc = abi_current_ws(&ws);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const unsigned long long total = (unsigned long long)G * G * G;
	const size_t words = (size_t)(total / 32);
	unsigned int* d_edges = nullptr;
	unsigned int* d_counts = nullptr;
	unsigned long long *d_offsets = nullptr, *d_starts = nullptr, *d_sums = nullptr;
	LongRun* d_long = nullptr;
	unsigned int* d_nlong = nullptr;
	unsigned short* d_out = nullptr;
	auto cleanup = [&] { for (void* p : {(void*)d_edges, (void*)d_counts, (void*)d_offsets, (void*)d_starts, (void*)d_sums, (void*)d_long, (void*)d_nlong}) if (p) cudaFree(p); };
	BV_CU(cudaMalloc(&d_edges, words * sizeof(unsigned int)));
	binvox_edges_kernel<<<dim3(G / 256, G / 256, G), kBvBlock, 0, st>>>(d_table, G, d_edges);
	g_launch_count++;
	BV_CU(cudaGetLastError());
	// run starts
	const size_t blocks = extract_blocks(words);
	BV_CU(cudaMalloc(&d_counts, (blocks + 1) * sizeof(unsigned int)));
	BV_CU(cudaMalloc(&d_offsets, (blocks + 1) * sizeof(unsigned long long)));
	BV_CU(launch_extract_count(d_edges, words, d_counts, d_offsets, st));
	unsigned long long n_runs = 0;
	BV_CU(cudaMemcpyAsync(&n_runs, d_offsets + blocks, sizeof(n_runs), cudaMemcpyDeviceToHost, st));
	BV_CU(cudaStreamSynchronize(st));
	BV_CU(cudaMalloc(&d_starts, (size_t)n_runs * sizeof(unsigned long long)));
	BV_CU(launch_extract_write(d_edges, words, d_offsets, 0ull, d_starts, st));
	// the first voxel's value: run 0 (voxel (0,0,0) is the top bit of word 0)
	unsigned int word0 = 0;
	BV_CU(cudaMemcpyAsync(&word0, d_table, sizeof(word0), cudaMemcpyDeviceToHost, st));
	// pairs per run -> offsets
	const size_t run_blocks = (size_t)((n_runs + kRunsPerBlock - 1) / kRunsPerBlock);
	BV_CU(cudaMalloc(&d_sums, (run_blocks + 1) * sizeof(unsigned long long)));
	binvox_count_kernel<<<(unsigned int)run_blocks, kBvBlock, 0, st>>>(d_starts, n_runs, total, d_sums);
	binvox_scan_kernel<<<1, 1024, 0, st>>>(d_sums, run_blocks);
	g_launch_count += 2;
	BV_CU(cudaGetLastError());
	unsigned long long n_pairs = 0;
	BV_CU(cudaMemcpyAsync(&n_pairs, d_sums + run_blocks, sizeof(n_pairs), cudaMemcpyDeviceToHost, st));
	BV_CU(cudaStreamSynchronize(st));
	const size_t max_long = (size_t)(total / (255ull * kLongRunPairs)) + 2;
	BV_CU(cudaMalloc(&d_long, max_long * sizeof(LongRun)));
	BV_CU(cudaMalloc(&d_nlong, sizeof(unsigned int)));
	BV_CU(cudaMemsetAsync(d_nlong, 0, sizeof(unsigned int), st));
	BV_CU(cudaMalloc(&d_out, (size_t)n_pairs * sizeof(unsigned short) + 16));
	binvox_emit_kernel<<<(unsigned int)run_blocks, kBvBlock, 0, st>>>(d_starts, n_runs, total, d_sums, word0 >> 31, d_out, d_long, d_nlong);
	binvox_long_kernel<<<(unsigned int)(ws->sm_count * 8), kBvBlock, 0, st>>>(d_long, d_nlong, d_out);
	g_launch_count += 2;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cleanup();
	if (e != cudaSuccess) { cudaFree(d_out); return abi_fail_cuda(e, "voxb200_binvox_rle"); }
	*d_bytes = reinterpret_cast<unsigned char*>(d_out);
	*n_bytes = (size_t)n_pairs * 2;
	return VOXB200_OK;
}
