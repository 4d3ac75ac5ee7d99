/*
 * voxelize_dropin.h — the reference's C++ host entry points, re-implemented on top of voxb200.h.
 *
 * An unmodified caller written against the reference (src/main.cpp:23-24 forward-declares exactly
 * these two functions; src/util_cuda.h:12 declares initCuda) links against libvoxb200.so without
 * source changes: same names, same C++ signatures (hence same mangled symbols), same synchronous
 * behaviour, same "[Perf] Voxelization GPU time" line, same print-and-exit(EXIT_FAILURE) on CUDA
 * errors (src/libs/cuda/helper_cuda.h:566-579), same OR / XOR *into* a caller-zeroed table.
 *
 * `voxinfo` / `AABox<T>` are layout-compatible restatements of src/util.h:41-69 (64 bytes:
 * bbox.min @0, bbox.max @12, gridsize @24, n_triangles @40, unit @48).  When the caller already
 * includes the reference's own util.h, define VOXB200_REFERENCE_TYPES_PROVIDED before this header.
 */
#pragma once
#include <cstddef>
#include <vector_types.h>

#ifndef VOXB200_REFERENCE_TYPES_PROVIDED
template <typename T>
struct AABox {
	T min;
	T max;
	AABox() : min(T()), max(T()) {}
	AABox(T lo, T hi) : min(lo), max(hi) {}
};

struct voxinfo {
	AABox<float3> bbox;
	uint3 gridsize;
	size_t n_triangles;
	float3 unit;
	// unit = bbox extent / gridsize per axis, in binary32 (src/util.h:56-61)
	voxinfo(const AABox<float3> box, const uint3 grid, const size_t triangles) : bbox(box), gridsize(grid), n_triangles(triangles) {
		unit.x = (box.max.x - box.min.x) / float(grid.x);
		unit.y = (box.max.y - box.min.y) / float(grid.y);
		unit.z = (box.max.z - box.min.z) / float(grid.z);
	}
};
#endif

// src/voxelize.cu:192 and src/voxelize_solid.cu:147.  triangle_data: 9 floats per triangle, device
// accessible; vtable: device accessible, zeroed by the caller (the reference never clears it).
void voxelize(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code);
void voxelize_solid(const voxinfo& v, float* triangle_data, unsigned int* vtable, bool morton_code);
// src/util_cuda.cpp:4-42: true when a usable CUDA device was selected.
bool initCuda();
