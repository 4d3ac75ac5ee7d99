// TEST INFRASTRUCTURE — not product code.
//
// Stand-in for the un-vendored trimesh2 dependency (github.com/Forceflow/trimesh2,
// tag 2022.03.04 per /root/reference/.github/workflows/autobuild.yml:20), which is not
// installed in this image.  It declares only what the reference's CPU voxelizer touches:
//   * trimesh::vec3 / trimesh::point with .x .y .z, a 3-float ctor and operator-
//     (cpu_voxelizer.cpp:40,44,246,250; util.h:11-18)
//   * trimesh::TriMesh::Face with operator[]   (cpu_voxelizer.cpp:63-65,256-258)
//   * trimesh::TriMesh {vertices, faces, bbox{min,max}}  (main.cpp:174-184)
// so that /root/reference/src/cpu_voxelizer.cpp compiles UNMODIFIED from where it lies.
// The only arithmetic here is one float subtraction per component.
#pragma once
#include <vector>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

namespace trimesh {

struct vec3 {
	float x, y, z;
	vec3() : x(0.0f), y(0.0f), z(0.0f) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
typedef vec3 point;

struct box3 {
	point min, max;
};

struct TriMesh {
	struct Face {
		int v[3];
		Face() { v[0] = v[1] = v[2] = 0; }
		Face(int a, int b, int c) { v[0] = a; v[1] = b; v[2] = c; }
		int& operator[](int i) { return v[i]; }
		const int& operator[](int i) const { return v[i]; }
	};
	std::vector<point> vertices;
	std::vector<Face> faces;
	box3 bbox;
	void need_faces() {}
	// min/max over ALL vertices, as trimesh2's need_bbox does
	void need_bbox() {
		if (vertices.empty()) return;
		bbox.min = bbox.max = vertices[0];
		for (size_t i = 1; i < vertices.size(); i++) {
			for (int k = 0; k < 3; k++) {
				if (vertices[i][k] < bbox.min[k]) bbox.min[k] = vertices[i][k];
				if (vertices[i][k] > bbox.max[k]) bbox.max[k] = vertices[i][k];
			}
		}
	}
	static void set_verbose(bool) {}
	// Minimal OBJ reader ("v x y z", "f a b c" with optional /t/n suffixes, polygons fanned) so that the reference's
	// UNMODIFIED main.cpp can load the test fixture (main.cpp:174).  write_obj_cubes' reorder round trip
	// (util_io.cpp:140-149) also lands here: it re-reads its own output and writes nothing back.
	static TriMesh* read(const char* path) {
		TriMesh* m = new TriMesh();
		FILE* f = std::fopen(path, "rb");
		if (!f) return m;
		char line[1 << 14];
		while (std::fgets(line, sizeof(line), f)) {
			if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
				char* q = line + 1;
				float x = std::strtof(q, &q), y = std::strtof(q, &q), z = std::strtof(q, &q);
				m->vertices.push_back(point(x, y, z));
			} else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
				std::vector<int> idx;
				char* q = line + 1;
				for (;;) {
					while (*q == ' ' || *q == '\t') q++;
					if (*q == 0 || *q == '\n' || *q == '\r') break;
					char* e;
					long v = std::strtol(q, &e, 10);
					if (e == q) break;
					idx.push_back(v > 0 ? (int)v - 1 : (int)m->vertices.size() + (int)v);
					q = e;
					while (*q && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') q++;
				}
				for (size_t k = 1; k + 1 < idx.size(); k++) m->faces.push_back(Face(idx[0], idx[k], idx[k + 1]));
			}
		}
		std::fclose(f);
		return m;
	}
	void write(const char*) {}
};

} // namespace trimesh
