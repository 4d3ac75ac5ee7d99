// extract.cu — device-side consumer of the bit table: compacts the set voxels into a list of voxel
// indices in ascending (table) order, so the host writers walk N_set voxels instead of calling checkVoxel G^3 times
// (util_io.cpp:92-285 does the latter) and only 8 bytes per SET voxel cross PCIe instead of the whole table.
//   extract_count_kernel   popcount per 8 KB block of the table
//   extract_scan_kernel    exclusive scan of the block counts (one CTA)
//   extract_write_kernel   per block: per-thread popcounts -> block scan -> every set bit written as its voxel index
#include "vox_internal.h"

namespace voxb {

constexpr int kExtBlock = 256;
constexpr int kExtWords = 8;                                   // words per thread
constexpr int kExtBlockWords = kExtBlock * kExtWords;          // 2048 words = 8 KB per block

__device__ __forceinline__ unsigned int block_sum(unsigned int v, unsigned int* smem, unsigned int& excl) {
	// inclusive warp scan, then scan of the 8 warp totals; returns the block total, excl = exclusive prefix of the caller
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned int inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned int up = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += up;
	}
	if (lane == 31) smem[wid] = inc;
	__syncthreads();
	unsigned int warp_base = 0, total = 0;
#pragma unroll
	for (int w = 0; w < kExtBlock / 32; w++) {
		const unsigned int s = smem[w];
		if (w < wid) warp_base += s;
		total += s;
	}
	excl = warp_base + inc - v;
	__syncthreads();
	return total;
}

__global__ void __launch_bounds__(kExtBlock) extract_count_kernel(const unsigned int* __restrict__ table, size_t n_words,
                                                                  unsigned int* __restrict__ block_counts) {
	__shared__ unsigned int smem[kExtBlock / 32];
	const size_t first = (size_t)blockIdx.x * kExtBlockWords + (size_t)threadIdx.x * kExtWords;
	unsigned int c = 0;
#pragma unroll
	for (int k = 0; k < kExtWords; k++) if (first + k < n_words) c += __popc(__ldg(table + first + k));
	unsigned int excl;
	const unsigned int total = block_sum(c, smem, excl);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// one CTA of 1024 threads: exclusive scan of n block counts (each < 2^17) into 64-bit offsets; offsets[n] = total.
// Rounds of 1024 consecutive counts — coalesced loads, the next round's value fetched before this round's scan — instead of one
// contiguous chunk per thread (whose strided 4-byte loads made the kernel latency-bound: 228 us for the 131,072 counts of a 1 GiB
// table, more than the pass that produced them).
__global__ void __launch_bounds__(1024) extract_scan_kernel(const unsigned int* __restrict__ counts, unsigned long long* __restrict__ offsets, size_t n) {
	__shared__ unsigned int warp_total[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned long long base = 0;
	unsigned int next = threadIdx.x < n ? counts[threadIdx.x] : 0u;
	for (size_t r = 0; r < n; r += 1024) {
		const unsigned int c = next;
		const size_t ahead = r + 1024 + threadIdx.x;
		next = ahead < n ? counts[ahead] : 0u;
		unsigned int inc = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned int up = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += up;
		}
		if (lane == 31) warp_total[wid] = inc;
		__syncthreads();
		unsigned int wt = warp_total[lane];                     // every warp scans the 32 warp totals itself
		unsigned int winc = wt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned int up = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc += up;
		}
		const unsigned int warp_base = __shfl_sync(0xffffffffu, winc - wt, wid);
		const unsigned int round_total = __shfl_sync(0xffffffffu, winc, 31);
		if (r + threadIdx.x < n) offsets[r + threadIdx.x] = base + warp_base + (inc - c);
		base += round_total;
		__syncthreads();
	}
	if (threadIdx.x == 0) offsets[n] = base;
}

__global__ void __launch_bounds__(kExtBlock) extract_write_kernel(const unsigned int* __restrict__ table, size_t n_words,
                                                                  const unsigned long long* __restrict__ offsets,
                                                                  unsigned long long first_voxel, unsigned long long* __restrict__ out) {
	__shared__ unsigned int smem[kExtBlock / 32];
	const size_t first = (size_t)blockIdx.x * kExtBlockWords + (size_t)threadIdx.x * kExtWords;
	unsigned int w[kExtWords], c = 0;
#pragma unroll
	for (int k = 0; k < kExtWords; k++) { w[k] = first + k < n_words ? __ldg(table + first + k) : 0u; c += __popc(w[k]); }
	unsigned int excl;
	block_sum(c, smem, excl);
	unsigned long long* o = out + offsets[blockIdx.x] + excl;
#pragma unroll
	for (int k = 0; k < kExtWords; k++) {
		unsigned int bits = w[k];
		const unsigned long long base = first_voxel + ((unsigned long long)(first + k) << 5);
		while (bits) {                                      // voxel idx%32 = 0 is the MSB: take the highest set bit first
			const int msb = 31 - __clz(bits);
			bits &= ~(1u << msb);
			*o++ = base + (unsigned long long)(31 - msb);
		}
	}
}

// ---- non-zero WORDS of the table as {word index, value} pairs, ascending: the sparse read-back of readback.cu ----------------
// Same three passes as above (count per 8 KB block, scan, write); a thread owns 8 consecutive words = two 16-byte loads.
__global__ void __launch_bounds__(kExtBlock) nz_count_kernel(const uint4* __restrict__ table, size_t n_words, unsigned int* __restrict__ block_counts) {
	__shared__ unsigned int smem[kExtBlock / 32];
	const size_t first = (size_t)blockIdx.x * kExtBlockWords + (size_t)threadIdx.x * kExtWords;
	unsigned int c = 0;
	if (first + kExtWords <= n_words) {
		const uint4 a = __ldg(table + first / 4), b = __ldg(table + first / 4 + 1);
		c = (a.x != 0u) + (a.y != 0u) + (a.z != 0u) + (a.w != 0u) + (b.x != 0u) + (b.y != 0u) + (b.z != 0u) + (b.w != 0u);
	} else {
		const unsigned int* t = reinterpret_cast<const unsigned int*>(table);
		for (int k = 0; k < kExtWords; k++) if (first + k < n_words) c += __ldg(t + first + k) != 0u;
	}
	unsigned int excl;
	const unsigned int total = block_sum(c, smem, excl);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kExtBlock) nz_write_kernel(const uint4* __restrict__ table, size_t n_words, const unsigned long long* __restrict__ offsets,
                                                             uint2* __restrict__ out) {
	__shared__ unsigned int smem[kExtBlock / 32];
	const size_t first = (size_t)blockIdx.x * kExtBlockWords + (size_t)threadIdx.x * kExtWords;
	unsigned int w[kExtWords], c = 0;
	if (first + kExtWords <= n_words) {
		const uint4 a = __ldg(table + first / 4), b = __ldg(table + first / 4 + 1);
		w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
	} else {
		const unsigned int* t = reinterpret_cast<const unsigned int*>(table);
#pragma unroll
		for (int k = 0; k < kExtWords; k++) w[k] = first + k < n_words ? __ldg(t + first + k) : 0u;
	}
#pragma unroll
	for (int k = 0; k < kExtWords; k++) c += w[k] != 0u;
	unsigned int excl;
	const unsigned int total = block_sum(c, smem, excl);
	if (total == 0u) return;
	uint2* o = out + offsets[blockIdx.x] + excl;
#pragma unroll
	for (int k = 0; k < kExtWords; k++) if (w[k]) *o++ = make_uint2((unsigned int)(first + k), w[k]);
}
cudaError_t launch_nz_count(const unsigned int* d_table, size_t n_words, unsigned int* d_counts, unsigned long long* d_offsets, cudaStream_t st) {
	const size_t blocks = (n_words + kExtBlockWords - 1) / kExtBlockWords;
	if (blocks == 0) return cudaSuccess;
	nz_count_kernel<<<(unsigned)blocks, kExtBlock, 0, st>>>(reinterpret_cast<const uint4*>(d_table), n_words, d_counts);
	extract_scan_kernel<<<1, 1024, 0, st>>>(d_counts, d_offsets, blocks);
	g_launch_count += 2;
	return cudaGetLastError();
}
cudaError_t launch_nz_write(const unsigned int* d_table, size_t n_words, const unsigned long long* d_offsets, void* d_pairs, cudaStream_t st) {
	const size_t blocks = (n_words + kExtBlockWords - 1) / kExtBlockWords;
	if (blocks == 0) return cudaSuccess;
	nz_write_kernel<<<(unsigned)blocks, kExtBlock, 0, st>>>(reinterpret_cast<const uint4*>(d_table), n_words, d_offsets, reinterpret_cast<uint2*>(d_pairs));
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_extract_count(const unsigned int* d_table, size_t n_words, unsigned int* d_counts, unsigned long long* d_offsets, cudaStream_t st) {
	const size_t blocks = (n_words + kExtBlockWords - 1) / kExtBlockWords;
	if (blocks == 0) return cudaSuccess;
	extract_count_kernel<<<(unsigned)blocks, kExtBlock, 0, st>>>(d_table, n_words, d_counts);
	extract_scan_kernel<<<1, 1024, 0, st>>>(d_counts, d_offsets, blocks);
	g_launch_count += 2;
	return cudaGetLastError();
}
cudaError_t launch_extract_write(const unsigned int* d_table, size_t n_words, const unsigned long long* d_offsets, unsigned long long first_voxel,
                                 unsigned long long* d_out, cudaStream_t st) {
	const size_t blocks = (n_words + kExtBlockWords - 1) / kExtBlockWords;
	if (blocks == 0) return cudaSuccess;
	extract_write_kernel<<<(unsigned)blocks, kExtBlock, 0, st>>>(d_table, n_words, d_offsets, first_voxel, d_out);
	g_launch_count++;
	return cudaGetLastError();
}
size_t extract_blocks(size_t n_words) { return (n_words + kExtBlockWords - 1) / kExtBlockWords; }

}  // namespace voxb
