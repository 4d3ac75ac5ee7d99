// TEST INFRASTRUCTURE — not product code.
// C-ABI driver around the reference's UNMODIFIED output writers (src/util_io.cpp + the vendored MagicaVoxel
// writer), compiled from where they lie by oracle/Makefile into oracle/_ref/libvoxref_io.so.  Used by
// tests/golden/make_golden.py to produce golden output files for the CLI's writers.
#include <cstring>
#include "util_io.h"

extern "C" int voxref_write(int format, const unsigned int* table, unsigned int gridsize, const float* bbox_min, const float* bbox_max,
                            size_t n_triangles, const char* base_filename) {
	AABox<float3> box(make_float3(bbox_min[0], bbox_min[1], bbox_min[2]), make_float3(bbox_max[0], bbox_max[1], bbox_max[2]));
	voxinfo info(box, make_uint3(gridsize, gridsize, gridsize), n_triangles);
	size_t bytes = static_cast<size_t>(ceil(static_cast<size_t>(gridsize) * static_cast<size_t>(gridsize) * static_cast<size_t>(gridsize) / 32.0f) * 4);
	switch (format) {
		case 0: write_binvox(table, info, base_filename); break;
		case 1: write_binary((void*)table, bytes, base_filename); break;
		case 2: write_obj_pointcloud(table, info, base_filename); break;
		case 3: write_obj_cubes(table, info, base_filename); break;
		case 4: write_vox(table, info, base_filename); break;
		default: return 1;
	}
	return 0;
}
