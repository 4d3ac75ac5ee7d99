// TEST INFRASTRUCTURE — stand-in for trimesh2's TriMesh_algo.h, which the reference's util_io.h includes.
// Only write_obj_cubes touches it (TriMesh::read -> reorder_verts -> write, util_io.cpp:140-149): the stubs below
// make that round trip a no-op, so the reference's obj-cubes output is its own raw cube file.
#pragma once
#include "TriMesh.h"
namespace trimesh {
inline void reorder_verts(TriMesh*) {}
}
