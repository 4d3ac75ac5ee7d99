// Micro-benchmark: after zero-filling X MB with plain 16-byte stores (or evict_last hinted stores), how expensive is one
// atomic OR per 32-byte sector over the same X MB?  (Are freshly written lines L2 hits for atomics?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_retention l2_retention.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void zero_plain(uint4* p, size_t n16) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}
__global__ void zero_hint(uint4* p, size_t n16) {
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
		asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(p + i), "r"(0u), "l"(pol) : "memory");
}
// one RED per sector, sectors visited in a scrambled order (like scattered triangles)
__global__ void red_sectors(unsigned int* p, size_t n_sectors, unsigned int stride) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sectors; i += (size_t)gridDim.x * blockDim.x) {
		const size_t s = (i * stride) % n_sectors;
		asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p + s * 8 + (i & 7)), "r"(1u << (i & 31)) : "memory");
	}
}
int main() {
	const size_t max_bytes = 1ull << 30;
	unsigned int* buf;
	cudaMalloc(&buf, max_bytes);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int mode = 0; mode < 3; mode++)            // 0: plain zero then RED, 1: hinted zero then RED, 2: RED only (table cold in DRAM)
		for (size_t mb = 8; mb <= 512; mb *= 2) {
			const size_t bytes = mb << 20, n16 = bytes / 16, n_sectors = bytes / 32;
			float best = 1e9f, bestz = 1e9f;
			for (int rep = 0; rep < 5; rep++) {
				if (mode == 2) { zero_plain<<<148 * 16, 512>>>(reinterpret_cast<uint4*>(buf), max_bytes / 16); }      // flush: the X MB end up evicted
				cudaEventRecord(e0);
				if (mode == 0) zero_plain<<<148 * 16, 512>>>(reinterpret_cast<uint4*>(buf), n16);
				if (mode == 1) zero_hint<<<148 * 16, 512>>>(reinterpret_cast<uint4*>(buf), n16);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float msz; cudaEventElapsedTime(&msz, e0, e1);
				cudaEventRecord(e0);
				red_sectors<<<148 * 16, 256>>>(buf, n_sectors, 7919u);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float ms; cudaEventElapsedTime(&ms, e0, e1);
				if (ms < best) best = ms;
				if (msz < bestz) bestz = msz;
			}
			printf("mode %d  %4zu MB: zero %.4f ms, %zu REDs (1 per sector) %.4f ms = %.1f G sectors/s\n", mode, mb, bestz, n_sectors, best, n_sectors / best * 1e-6);
		}
	return 0;
}
