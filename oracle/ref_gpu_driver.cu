// TEST / BENCH INFRASTRUCTURE — not product code.  Only tests/ and bench.py's `ref_gpu_baseline` leg may load the
// library built from this file.
//
// C-ABI driver around the reference's UNMODIFIED GPU kernels: compiled together with /root/reference/src/voxelize.cu
// and voxelize_solid.cu (read in place, never copied) by oracle/Makefile into oracle/_ref/libvoxref_gpu.so, for
// sm_100a.  It is the second baseline of bench.py: the reference's own one-thread-per-triangle kernels
// (voxelize.cu:58-190, voxelize_solid.cu:73-145) launched by the reference's own host entry points
// (voxelize.cu:192-238, voxelize_solid.cu:147-193) on the same B200 as the product, with the triangles and the
// table resident in device memory (plain cudaMalloc: what the reference's managed buffers are once they have been
// faulted in), timed with CUDA events around the call.  These kernels are NOT the parity target (they differ from
// the reference's CPU path by construction): the table they produce is only popcounted.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <unistd.h>
#include <fcntl.h>
#include <cuda_runtime.h>
#include "util.h"

// main.cpp:23-24
void voxelize(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code);
void voxelize_solid(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code);

namespace {
struct StdoutMute {      // the reference prints "[Perf] Voxelization GPU time" per call
	int saved;
	StdoutMute() {
		fflush(stdout);
		saved = dup(1);
		int nul = open("/dev/null", O_WRONLY);
		if (nul >= 0) { dup2(nul, 1); close(nul); }
	}
	~StdoutMute() {
		fflush(stdout);
		if (saved >= 0) { dup2(saved, 1); close(saved); }
	}
};
}

extern "C" {

// bbox6 = the cubed, padded mesh bbox (min xyz, max xyz) exactly as the product's voxb200_make_grid returns it.
// Runs `warmup` untimed and `reps` timed calls; every call starts from a zeroed table (cudaMemsetAsync, timed apart).
// out_ms[0] = mean ms of the reference call (kernel + its own synchronisation), out_ms[1] = best, out_ms[2] = mean ms of the
// table memset the reference leaves to its caller.  host_table (optional) receives the table of the last call.
// Returns 0, or the CUDA error code.
int voxrefgpu_run(const float* bbox6, unsigned int gridsize, const float* host_tris9, size_t n_tris, int solid, int morton,
                  int warmup, int reps, float* out_ms, unsigned int* host_table) {
	AABox<float3> cube(make_float3(bbox6[0], bbox6[1], bbox6[2]), make_float3(bbox6[3], bbox6[4], bbox6[5]));
	voxinfo info(cube, make_uint3(gridsize, gridsize, gridsize), n_tris);
	const size_t G = gridsize;
	const size_t vtable_size = static_cast<size_t>(ceil(G * G * G / 32.0f) * 4);        // main.cpp:190
	float* d_tris = nullptr;
	unsigned int* d_table = nullptr;
	cudaError_t e = cudaMalloc(&d_tris, n_tris * 9 * sizeof(float) + 16);
	if (e == cudaSuccess) e = cudaMalloc(&d_table, vtable_size + 16);
	if (e == cudaSuccess) e = cudaMemcpy(d_tris, host_tris9, n_tris * 9 * sizeof(float), cudaMemcpyHostToDevice);
	cudaEvent_t ev[3];
	for (auto& x : ev) cudaEventCreate(&x);
	double sum = 0.0, sum_set = 0.0;
	float best = 1e30f;
	if (e == cudaSuccess) {
		StdoutMute mute;
		for (int i = 0; i < warmup + reps; i++) {
			cudaEventRecord(ev[0], 0);
			cudaMemsetAsync(d_table, 0, vtable_size, 0);
			cudaEventRecord(ev[1], 0);
			if (solid) voxelize_solid(info, d_tris, d_table, morton != 0);
			else voxelize(info, d_tris, d_table, morton != 0);
			cudaEventRecord(ev[2], 0);
			e = cudaEventSynchronize(ev[2]);
			if (e != cudaSuccess) break;
			float ms_set = 0.f, ms = 0.f;
			cudaEventElapsedTime(&ms_set, ev[0], ev[1]);
			cudaEventElapsedTime(&ms, ev[1], ev[2]);
			if (i >= warmup) { sum += ms; sum_set += ms_set; if (ms < best) best = ms; }
		}
	}
	if (e == cudaSuccess && host_table) e = cudaMemcpy(host_table, d_table, vtable_size, cudaMemcpyDeviceToHost);
	if (out_ms && reps > 0) { out_ms[0] = (float)(sum / reps); out_ms[1] = best; out_ms[2] = (float)(sum_set / reps); }
	for (auto& x : ev) cudaEventDestroy(x);
	cudaFree(d_tris);
	cudaFree(d_table);
	return (int)e;
}

}  // extern "C"
