// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
//
// C-ABI driver around the reference's UNMODIFIED CPU voxelizer.  It is compiled together
// with /root/reference/src/cpu_voxelizer.cpp (read in place, never copied) by oracle/Makefile
// into oracle/_ref/libvoxref.so.  It reproduces what the reference's main() does between
// "mesh loaded" and "table filled" (main.cpp:179-190, 229-235):
//   need_bbox -> createMeshBBCube (util.h:80-110) -> voxinfo (util.h:50-61) -> calloc'd table
//   -> cpu_voxelizer::cpu_voxelize_mesh{,_solid} (cpu_voxelizer.cpp:35,241).
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <unistd.h>
#include <fcntl.h>
#include <time.h>
#include "cpu_voxelizer.h"

namespace {
// The reference prints "[Info] Using %d threads" / "[Perf] ..." on stdout from inside the
// voxelizer; mute fd 1 around the call so test and bench output stays machine-readable.
struct StdoutMute {
	int saved;
	explicit StdoutMute(bool on) : saved(-1) {
		if (!on) return;
		fflush(stdout);
		saved = dup(1);
		int nul = open("/dev/null", O_WRONLY);
		if (nul >= 0) { dup2(nul, 1); close(nul); }
	}
	~StdoutMute() {
		if (saved < 0) return;
		fflush(stdout);
		dup2(saved, 1);
		close(saved);
	}
};
}

extern "C" {

// Fills out[0..15] with the voxinfo the reference would build for this mesh and grid size:
// bbox.min xyz, bbox.max xyz, unit xyz (9 floats), then 0-padding.  Returns sizeof(voxinfo).
int voxref_voxinfo(const float* verts, size_t nv, unsigned int gridsize, size_t n_tris, float* out) {
	trimesh::TriMesh mesh;
	mesh.vertices.resize(nv);
	for (size_t i = 0; i < nv; i++) mesh.vertices[i] = trimesh::point(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
	mesh.need_bbox();
	AABox<float3> cube = createMeshBBCube<float3>(AABox<float3>(trimesh_to_float3(mesh.bbox.min), trimesh_to_float3(mesh.bbox.max)));
	voxinfo info(cube, make_uint3(gridsize, gridsize, gridsize), n_tris);
	out[0] = info.bbox.min.x; out[1] = info.bbox.min.y; out[2] = info.bbox.min.z;
	out[3] = info.bbox.max.x; out[4] = info.bbox.max.y; out[5] = info.bbox.max.z;
	out[6] = info.unit.x; out[7] = info.unit.y; out[8] = info.unit.z;
	return (int)sizeof(voxinfo);
}

// Field offsets of the reference's voxinfo, for the ABI test of include/voxinfo_compat.h.
void voxref_voxinfo_layout(size_t* out) {
	out[0] = sizeof(voxinfo);
	out[1] = offsetof(voxinfo, bbox);
	out[2] = offsetof(voxinfo, gridsize);
	out[3] = offsetof(voxinfo, n_triangles);
	out[4] = offsetof(voxinfo, unit);
	out[5] = alignof(voxinfo);
}

// Runs the reference CPU path on an indexed mesh.  table must hold ceil(G^3/32)*4 bytes; it is
// zeroed here (the reference callocs it, main.cpp:229).  Returns the reference-equivalent
// elapsed milliseconds (wall clock around the call, which spans the same vertex-shift prepass +
// triangle loop as the reference's own Timer, cpu_voxelizer.cpp:36,177).
double voxref_voxelize(const float* verts, size_t nv, const int32_t* faces, size_t nf,
                       unsigned int gridsize, int solid, int morton, unsigned int* table, int quiet) {
	trimesh::TriMesh mesh;
	mesh.vertices.resize(nv);
	for (size_t i = 0; i < nv; i++) mesh.vertices[i] = trimesh::point(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
	mesh.faces.resize(nf);
	for (size_t i = 0; i < nf; i++) mesh.faces[i] = trimesh::TriMesh::Face(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]);
	mesh.need_bbox();
	AABox<float3> cube = createMeshBBCube<float3>(AABox<float3>(trimesh_to_float3(mesh.bbox.min), trimesh_to_float3(mesh.bbox.max)));
	voxinfo info(cube, make_uint3(gridsize, gridsize, gridsize), mesh.faces.size());
	size_t vtable_size = static_cast<size_t>(ceil(static_cast<size_t>(gridsize) * static_cast<size_t>(gridsize) * static_cast<size_t>(gridsize) / 32.0f) * 4);
	memset(table, 0, vtable_size);
	StdoutMute mute(quiet != 0);
	timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	if (solid) cpu_voxelizer::cpu_voxelize_mesh_solid(info, &mesh, table, morton != 0);
	else cpu_voxelizer::cpu_voxelize_mesh(info, &mesh, table, morton != 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
}

int voxref_max_threads(void) { return omp_get_max_threads(); }
void voxref_set_threads(int n) { omp_set_num_threads(n); }

} // extern "C"
