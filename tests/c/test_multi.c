/* C caller of the multi-GPU entry point (include/voxb200.h), the way the reference's main() would use it (main.cpp:203-222):
 * indexed mesh and table in host memory, no Python, no torch, no NCCL.
 *   test_multi <mesh.bin> <gridsize> <flags> <n_devices>
 * mesh.bin: u64 n_verts, u64 n_faces, float verts[3 n_verts], int32 faces[3 n_faces].
 * Prints "fnv1a64 <hex> popcount <n> devices <n> total_ms <ms>"; the test compares the hash with the oracle's. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "voxb200.h"

static void die(const char* what, int rc) {
	fprintf(stderr, "%s failed: %d: %s\n", what, rc, voxb200_last_error());
	exit(1);
}

int main(int argc, char** argv) {
	if (argc < 5) { fprintf(stderr, "usage: test_multi mesh.bin gridsize flags n_devices\n"); return 2; }
	FILE* f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	uint64_t nv = 0, nf = 0;
	if (fread(&nv, 8, 1, f) != 1 || fread(&nf, 8, 1, f) != 1) return 2;
	const unsigned int G = (unsigned int)atoi(argv[2]), flags = (unsigned int)atoi(argv[3]);
	int n_devices = atoi(argv[4]), have = 0, rc;
	if ((rc = voxb200_device_count(&have))) die("voxb200_device_count", rc);
	if (n_devices <= 0 || n_devices > have) n_devices = have;
	float* verts; int32_t* faces; unsigned int* table;
	const size_t table_bytes = voxb200_table_bytes(G);
	if ((rc = voxb200_host_alloc((void**)&verts, nv * 12))) die("voxb200_host_alloc", rc);
	if ((rc = voxb200_host_alloc((void**)&faces, nf * 12))) die("voxb200_host_alloc", rc);
	if ((rc = voxb200_host_alloc((void**)&table, table_bytes))) die("voxb200_host_alloc", rc);
	if (fread(verts, 12, nv, f) != nv || fread(faces, 12, nf, f) != nf) return 2;
	fclose(f);
	float mn[3] = {verts[0], verts[1], verts[2]}, mx[3] = {verts[0], verts[1], verts[2]};
	for (uint64_t i = 0; i < nv; i++)
		for (int k = 0; k < 3; k++) {
			const float v = verts[3 * i + k];
			if (v < mn[k]) mn[k] = v;
			if (v > mx[k]) mx[k] = v;
		}
	voxb200_grid grid;
	if ((rc = voxb200_make_grid(mn, mx, G, nf, &grid))) die("voxb200_make_grid", rc);
	float ms[8];
	memset(table, 0xff, table_bytes);                 /* every byte must be overwritten */
	for (int rep = 0; rep < 2; rep++)                  /* second call: buffers and prepared meshes are reused */
		if ((rc = voxb200_voxelize_host_multi(&grid, verts, nv, faces, table, flags, NULL, n_devices, ms))) die("voxb200_voxelize_host_multi", rc);
	uint64_t h = 0xcbf29ce484222325ull, pop = 0;
	const unsigned char* b = (const unsigned char*)table;
	for (size_t i = 0; i < table_bytes; i++) { h = (h ^ b[i]) * 0x100000001b3ull; pop += (uint64_t)__builtin_popcount(b[i]); }
	printf("fnv1a64 %016llx popcount %llu devices %d total_ms %.3f\n", (unsigned long long)h, (unsigned long long)pop, n_devices, ms[5]);
	voxb200_host_free(verts); voxb200_host_free(faces); voxb200_host_free(table);
	voxb200_release();
	return 0;
}
