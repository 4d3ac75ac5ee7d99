// dropin.cu — C++ symbols of the reference's host API (include/voxelize_dropin.h), thin wrappers
// over the C ABI.  Behavioural contract copied from the reference launchers (voxelize.cu:192-238,
// voxelize_solid.cu:147-193): synchronous, times the device work with CUDA events, prints the
// "[Perf]" line, and turns any failure into print + exit(EXIT_FAILURE).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstddef>

#include "../../include/voxb200.h"
#include "../../include/voxelize_dropin.h"

static_assert(sizeof(voxinfo) == sizeof(voxb200_grid), "voxinfo / voxb200_grid size mismatch");
static_assert(offsetof(voxinfo, bbox) == offsetof(voxb200_grid, bbox_min), "bbox offset");
static_assert(offsetof(voxinfo, gridsize) == offsetof(voxb200_grid, gridsize), "gridsize offset");
static_assert(offsetof(voxinfo, n_triangles) == offsetof(voxb200_grid, n_triangles), "n_triangles offset");
static_assert(offsetof(voxinfo, unit) == offsetof(voxb200_grid, unit), "unit offset");
static_assert(sizeof(voxinfo) == 64 && offsetof(voxinfo, gridsize) == 24 && offsetof(voxinfo, n_triangles) == 40 && offsetof(voxinfo, unit) == 48,
              "layout must match the reference's voxinfo (util.h:50-69)");

namespace {
[[noreturn]] void die(const char* where, int rc) {
	fprintf(stderr, "CUDA error at %s code=%d \"%s\" \n", where, rc, voxb200_last_error());
	exit(EXIT_FAILURE);
}

void run(bool solid, const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code) {
	cudaEvent_t start_vox, stop_vox;
	if (cudaEventCreate(&start_vox) != cudaSuccess || cudaEventCreate(&stop_vox) != cudaSuccess) die("cudaEventCreate", VOXB200_ECUDA);
	const unsigned int flags = (morton_code ? VOXB200_MORTON : 0u) | VOXB200_ACCUMULATE;
	const voxb200_grid* grid = reinterpret_cast<const voxb200_grid*>(&v);
	// The caller sized vtable with the reference's formula (main.cpp:190), which is a word short of G^3 bits for some odd
	// grid sizes; the reference's kernels then write past it.  Here such a grid is voxelized in a library-owned table of the
	// full size and the caller's bytes are copied back, so nothing is written outside the caller's allocation.
	const size_t have = voxb200_reference_table_bytes(v.gridsize.x), need = voxb200_table_bytes(v.gridsize.x);
	unsigned int* work = vtable;
	if (need > have) {
		void* p = nullptr;
		if (int rc = voxb200_malloc(&p, need)) die("voxb200_malloc", rc);
		work = static_cast<unsigned int*>(p);
		if (cudaMemset(work, 0, need) != cudaSuccess || cudaMemcpy(work, vtable, have, cudaMemcpyDefault) != cudaSuccess) die("cudaMemcpy", VOXB200_ECUDA);
	}
	cudaEventRecord(start_vox, 0);
	const int rc = solid ? voxb200_solid(grid, triangle_data, work, flags, nullptr, nullptr)
	                     : voxb200_surface(grid, triangle_data, work, flags, nullptr, nullptr);
	if (rc) die(solid ? "voxelize_solid" : "voxelize", rc);
	cudaEventRecord(stop_vox, 0);
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess && work != vtable) {
		e = cudaMemcpy(vtable, work, have, cudaMemcpyDefault);
		voxb200_free(work);
	}
	if (e != cudaSuccess) { fprintf(stderr, "CUDA error at voxelize: code=%d(%s) \n", (int)e, cudaGetErrorName(e)); exit(EXIT_FAILURE); }
	uint64_t counters[4] = {0, 0, 0, 0};
	if (voxb200_last_counters(counters) != VOXB200_OK) die("voxb200_last_counters", VOXB200_ECUDA);
	if (counters[1] == ~0ull) {
		fprintf(stderr, "voxelize: more than 2^32 (y,z) rows queued for the large-triangle path; the table is not valid \n");
		exit(EXIT_FAILURE);
	}
	float elapsed = 0.0f;
	cudaEventElapsedTime(&elapsed, start_vox, stop_vox);
	printf("[Perf] Voxelization GPU time: %.1f ms\n", elapsed);
	cudaEventDestroy(start_vox);
	cudaEventDestroy(stop_vox);
}
}  // namespace

void voxelize(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code) { run(false, v, triangle_data, vtable, morton_code); }
void voxelize_solid(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code) { run(true, v, triangle_data, vtable, morton_code); }

bool initCuda() {
	int n = 0;
	if (voxb200_device_count(&n) != VOXB200_OK) {
		fprintf(stderr, "[CUDA] First call to CUDA Runtime API failed. Are the drivers installed? \n");
		return false;
	}
	if (n < 1) {
		fprintf(stderr, "[CUDA] No CUDA devices found. Make sure CUDA device is powered, connected and available. \n");
		return false;
	}
	// the reference picks the max-GFLOPS device (findCudaDevice); on an HGX B200 box all are equal: take 0
	if (voxb200_init(0) != VOXB200_OK) {
		fprintf(stderr, "[CUDA] %s \n", voxb200_last_error());
		return false;
	}
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	fprintf(stdout, "[CUDA] Best device: %s \n", prop.name);
	fprintf(stdout, "[CUDA] Available device memory: %llu of %llu MB \n", (unsigned long long)(free_b >> 20), (unsigned long long)(total_b >> 20));
	return true;
}
