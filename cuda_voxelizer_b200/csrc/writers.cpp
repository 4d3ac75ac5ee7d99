// writers.cpp — the five output formats of the reference CLI (src/util_io.cpp), restated for the drop-in tool.
// They consume the bit table on the host through check_voxel(), i.e. they depend only on the layout contract.
// File names follow the reference exactly (they are appended to the full input file name, main.cpp:246-257).
//   morton ...... <file>.bin                 raw table dump                       (util_io.cpp:192-200)
//   binvox ...... <file>_<G>.binvox          header + RLE, x -> z -> y order      (util_io.cpp:202-246)
//   obj_points .. <file>_<G>_pointcloud.obj  one "v" per set voxel centre         (util_io.cpp:154-190)
//   obj ......... <file>_<G>_voxels.obj      one cube (8 v, 12 f) per set voxel   (util_io.cpp:92-152; the reference
//                                            then round-trips the file through trimesh2's reorder_verts, which is
//                                            not available here: this writer stops at the raw cube mesh)
//   vox ......... <file>_<G>.vox             MagicaVoxel scene, same axis mapping (x, G - z, y) as util_io.cpp:276
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "cli.h"

namespace voxcli {

void write_binary(const void* data, size_t bytes, const std::string& base_filename) {
	const std::string name = base_filename + ".bin";
	fprintf(stdout, "[I/O] Writing data in binary format to %s (%zu bytes) \n", name.c_str(), bytes);
	std::ofstream out(name.c_str(), std::ios_base::out | std::ios_base::binary);
	out.write(static_cast<const char*>(data), (std::streamsize)bytes);
}

void write_binvox(const unsigned int* vtable, const voxinfo& info, const std::string& base_filename) {
	const std::string name = base_filename + "_" + std::to_string(info.gridsize.x) + ".binvox";
	fprintf(stdout, "[I/O] Writing data in binvox format to %s \n", name.c_str());
	std::ofstream out(name.c_str(), std::ios::out | std::ios::binary);
	const float sx = info.bbox.max.x - info.bbox.min.x, sy = info.bbox.max.y - info.bbox.min.y, sz = info.bbox.max.z - info.bbox.min.z;
	out << "#binvox 1" << std::endl;
	out << "dim " << info.gridsize.x << " " << info.gridsize.y << " " << info.gridsize.z << std::endl;
	out << "translate " << info.bbox.min.x << " " << info.bbox.min.y << " " << info.bbox.min.z << std::endl;
	out << "scale " << std::max(std::max(sx, sy), sz) << std::endl;
	out << "data" << std::endl;
	// run-length pairs (value, count<=255), voxels visited x-major, then z, then y
	std::vector<char> buf;
	buf.reserve(1 << 20);
	char value = 0;
	unsigned char run = 0;
	bool first = true;
	for (size_t x = 0; x < info.gridsize.x; x++)
		for (size_t z = 0; z < info.gridsize.z; z++)
			for (size_t y = 0; y < info.gridsize.y; y++) {
				const char v = check_voxel(x, y, z, info.gridsize, vtable) ? 1 : 0;
				if (first) { value = v; buf.push_back(value); run = 1; first = false; continue; }
				if (v != value || run == 255) {
					buf.push_back((char)run);
					run = 1;
					value = v;
					buf.push_back(value);
				} else {
					run++;
				}
				if (buf.size() >= (1u << 20) - 4) { out.write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
			}
	buf.push_back((char)run);
	out.write(buf.data(), (std::streamsize)buf.size());
}

void write_obj_pointcloud(const unsigned int* vtable, const voxinfo& info, const std::string& base_filename) {
	const std::string name = base_filename + "_" + std::to_string(info.gridsize.x) + "_pointcloud.obj";
	fprintf(stdout, "[I/O] Writing data in obj point cloud format to %s \n", name.c_str());
	std::ofstream out(name.c_str(), std::ios::out);
	for (size_t x = 0; x < info.gridsize.x; x++)
		for (size_t y = 0; y < info.gridsize.y; y++)
			for (size_t z = 0; z < info.gridsize.z; z++)
				if (check_voxel(x, y, z, info.gridsize, vtable)) out << "v " << (x + 0.5) << " " << (y + 0.5) << " " << (z + 0.5) << "\n";
}

void write_obj_cubes(const unsigned int* vtable, const voxinfo& info, const std::string& base_filename) {
	const std::string name = base_filename + "_" + std::to_string(info.gridsize.x) + "_voxels.obj";
	fprintf(stdout, "[I/O] Writing data in obj voxels format to file %s \n", name.c_str());
	std::ofstream out(name.c_str(), std::ios::out);
	// corner order and relative (negative) face indices as in util_io.cpp:45-89: v8 is written first, so corner i is -i
	static const int corner[8][3] = {{1, 1, 0}, {0, 1, 0}, {1, 0, 0}, {0, 0, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};   // v8..v1
	static const int face[12][3] = {{-1, -3, -4}, {-1, -4, -2}, {-4, -3, -6}, {-4, -6, -5}, {-3, -1, -8}, {-3, -8, -6},
	                                {-1, -2, -7}, {-1, -7, -8}, {-2, -4, -5}, {-2, -5, -7}, {-5, -6, -8}, {-5, -8, -7}};
	for (size_t x = 0; x < info.gridsize.x; x++)
		for (size_t y = 0; y < info.gridsize.y; y++)
			for (size_t z = 0; z < info.gridsize.z; z++) {
				if (!check_voxel(x, y, z, info.gridsize, vtable)) continue;
				for (const auto& c : corner) out << "v " << (long)x + c[0] << " " << (long)y + c[1] << " " << (long)z + c[2] << "\n";
				for (const auto& f : face) out << "f " << f[0] << " " << f[1] << " " << f[2] << "\n";
			}
}

// ---- MagicaVoxel .vox (format 150): models of at most 256^3 placed by a transform/group/shape scene graph
namespace {
void put32(std::vector<char>& b, int32_t v) { char t[4]; memcpy(t, &v, 4); b.insert(b.end(), t, t + 4); }
void put_str(std::vector<char>& b, const std::string& s) { put32(b, (int32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); }
void put_chunk(std::vector<char>& out, const char id[4], const std::vector<char>& content, const std::vector<char>& children = {}) {
	out.insert(out.end(), id, id + 4);
	put32(out, (int32_t)content.size());
	put32(out, (int32_t)children.size());
	out.insert(out.end(), content.begin(), content.end());
	out.insert(out.end(), children.begin(), children.end());
}
}  // namespace

void write_vox(const unsigned int* vtable, const voxinfo& info, const std::string& base_filename) {
	const std::string name = base_filename + "_" + std::to_string(info.gridsize.x) + ".vox";
	fprintf(stdout, "[I/O] Writing data in vox format to %s \n", name.c_str());
	const int kModel = 256;
	struct Key { int mx, my, mz; bool operator<(const Key& o) const { return mx != o.mx ? mx < o.mx : (my != o.my ? my < o.my : mz < o.mz); } };
	std::map<Key, std::vector<unsigned char>> models;     // xyzi quadruples
	const int G = (int)info.gridsize.x;
	for (int x = 0; x < G; x++)
		for (int y = 0; y < (int)info.gridsize.z; y++)
			for (int z = 0; z < (int)info.gridsize.y; z++) {
				if (!check_voxel(x, y, z, info.gridsize, vtable)) continue;
				const int vx = x, vy = -z + (int)info.gridsize.z, vz = y;   // the reference's axis mapping
				auto& m = models[Key{vx / kModel, vy / kModel, vz / kModel}];
				m.push_back((unsigned char)(vx % kModel)); m.push_back((unsigned char)(vy % kModel)); m.push_back((unsigned char)(vz % kModel)); m.push_back(1);
			}
	std::vector<char> children;
	for (auto& kv : models) {
		std::vector<char> size, xyzi;
		put32(size, kModel); put32(size, kModel); put32(size, kModel);
		put_chunk(children, "SIZE", size);
		put32(xyzi, (int32_t)(kv.second.size() / 4));
		xyzi.insert(xyzi.end(), kv.second.begin(), kv.second.end());
		put_chunk(children, "XYZI", xyzi);
	}
	if (models.size() > 1) {
		std::vector<char> c;
		put32(c, 0); put32(c, 0); put32(c, 1); put32(c, -1); put32(c, -1); put32(c, 1); put32(c, 0);     // root nTRN -> group 1
		put_chunk(children, "nTRN", c);
		c.clear();
		put32(c, 1); put32(c, 0); put32(c, (int32_t)models.size());
		for (size_t i = 0; i < models.size(); i++) put32(c, (int32_t)(2 + 2 * i));
		put_chunk(children, "nGRP", c);
		size_t i = 0;
		for (auto& kv : models) {
			c.clear();
			put32(c, (int32_t)(2 + 2 * i)); put32(c, 0); put32(c, (int32_t)(3 + 2 * i)); put32(c, -1); put32(c, 0); put32(c, 1);
			put32(c, 1);                                                                               // frame dict: 1 pair
			put_str(c, "_t");
			put_str(c, std::to_string(kv.first.mx * kModel + kModel / 2) + " " + std::to_string(kv.first.my * kModel + kModel / 2) + " " + std::to_string(kv.first.mz * kModel + kModel / 2));
			put_chunk(children, "nTRN", c);
			c.clear();
			put32(c, (int32_t)(3 + 2 * i)); put32(c, 0); put32(c, 1); put32(c, (int32_t)i); put32(c, 0);
			put_chunk(children, "nSHP", c);
			i++;
		}
	}
	std::vector<char> rgba(256 * 4, (char)255);
	put_chunk(children, "RGBA", rgba);
	std::vector<char> file = {'V', 'O', 'X', ' '};
	put32(file, 150);
	put_chunk(file, "MAIN", {}, children);
	std::ofstream out(name.c_str(), std::ios::out | std::ios::binary);
	out.write(file.data(), (std::streamsize)file.size());
}

}  // namespace voxcli
