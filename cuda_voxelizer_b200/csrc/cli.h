// cli.h — host side of the cuda_voxelizer command-line tool (the reference's src/main.cpp flag surface),
// written over the C ABI / drop-in symbols.  trimesh2 is not available, so mesh loading is local.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/voxelize_dropin.h"

namespace voxcli {

struct Mesh {
	std::vector<float> vertices;   // xyz per vertex
	std::vector<int32_t> faces;    // 3 vertex indices per triangle
	float bbox_min[3];
	float bbox_max[3];             // over ALL vertices (trimesh2 need_bbox semantics, main.cpp:179)
	size_t n_vertices() const { return vertices.size() / 3; }
	size_t n_faces() const { return faces.size() / 3; }
};

// OBJ ("v", "f" with a, a/t, a//n, a/t/n; negative indices), PLY (ASCII, binary little- / big-endian, triangle strips), OFF,
// STL (binary / ASCII) and 3DS; polygons through trimesh2's tess() rule (quads along the shorter diagonal).
bool load_mesh(const std::string& path, Mesh& mesh, std::string& error);

// The reader side of the table layout (util.h:25-38).
inline bool check_voxel(size_t x, size_t y, size_t z, const uint3 gridsize, const unsigned int* vtable) {
	const size_t location = x + (y * gridsize.x) + (z * gridsize.x * gridsize.y);
	return (vtable[location / 32] >> (31 - (location % 32))) & 1u;
}

// The set voxels of a (linear-order) table as ascending voxel indices x + G*y + G*G*z — what voxb200_extract_voxels returns.
struct VoxelList {
	std::vector<uint64_t> indices;
	unsigned int gridsize;
};

// Output formats of util_io.cpp, same file names (note they append to the FULL input file name, main.cpp:246-257) and
// the same bytes, produced from the voxel list instead of G^3 checkVoxel() calls.
void write_binary(const void* data, size_t bytes, const std::string& base_filename);                       // -o morton
void write_binvox(const VoxelList& vox, const voxinfo& info, const std::string& base_filename);             // -o binvox
void write_binvox_payload(const unsigned char* payload, size_t bytes, const voxinfo& info, const std::string& base_filename);   // -o binvox, G % 256 == 0
void write_obj_pointcloud(const VoxelList& vox, const voxinfo& info, const std::string& base_filename);     // -o obj_points
void write_obj_cubes(const VoxelList& vox, const voxinfo& info, const std::string& base_filename);          // -o obj
void write_vox(const VoxelList& vox, const voxinfo& info, const std::string& base_filename);                // -o vox

}  // namespace voxcli
