// cli_main.cpp — `cuda_voxelizer`: the reference's command-line surface (src/main.cpp:109-155 flags
// -f -s -o -cpu -solid -h, same defaults, same log sections) driving the B200 path through the C ABI.
//   -cpu is accepted but refused: this product has no CPU voxelizer (the reference's lives in oracle/ as a checker).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/voxb200.h"
#include "cli.h"

using namespace voxcli;

namespace {
enum class Format { binvox, morton, obj_points, obj_cubes, vox };
const char* kFormatNames[] = {"binvox file", "morton encoded blob", "obj file (pointcloud)", "obj file (cubes)", "magicavoxel file"};

struct Options {
	std::string filename;
	unsigned int gridsize = 256;
	Format format = Format::vox;
	bool force_cpu = false;
	bool solid = false;
	std::string from_table;    // test hook: skip the GPU, read a raw linear bit table from this file and only run the writer
	std::string dump_mesh;     // test hook: write the loaded mesh (u64 nv, u64 nf, float32 xyz[nv], int32 abc[nf]) to this file and exit
};

void print_help() {
	printf("\n## HELP  \n");
	printf("Program options: \n\n");
	printf(" -f <path to model file: .ply, .obj, .3ds> (required)\n");
	printf(" -s <voxelization grid size, power of 2: 8 -> 512, 1024, ... (default: 256)>\n");
	printf(" -o <output format: vox, binvox, obj, obj_points or morton (default: vox)>\n");
	printf(" -cpu : (reference flag) not available: this build voxelizes on a B200 only\n");
	printf(" -solid : Force solid voxelization (experimental, needs watertight model)\n\n");
	printf("Example: cuda_voxelizer -f /home/jeroen/bunny.ply -s 512\n\n");
}

Options parse(int argc, char* argv[]) {
	Options o;
	if (argc < 2) { printf("Not enough program parameters. \n \n"); print_help(); exit(0); }
	bool have_file = false;
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		if (a == "-f" && i + 1 < argc) {
			o.filename = argv[++i];
			have_file = true;
			FILE* probe = fopen(o.filename.c_str(), "rb");
			if (!probe) { printf("[Err] File does not exist / cannot access: %s \n", o.filename.c_str()); exit(1); }
			fclose(probe);
		} else if (a == "-s" && i + 1 < argc) {
			o.gridsize = (unsigned int)atoi(argv[++i]);
		} else if (a == "-h") {
			print_help(); exit(0);
		} else if (a == "-o" && i + 1 < argc) {
			std::string f = argv[++i];
			std::transform(f.begin(), f.end(), f.begin(), ::tolower);
			if (f == "binvox") o.format = Format::binvox;
			else if (f == "morton") o.format = Format::morton;
			else if (f == "obj") o.format = Format::obj_cubes;
			else if (f == "obj_points") o.format = Format::obj_points;
			else if (f == "vox") o.format = Format::vox;
			else { printf("[Err] Unrecognized output format: %s, valid options are binvox (default), morton, obj or obj_points \n", f.c_str()); exit(1); }
		} else if (a == "-cpu") {
			o.force_cpu = true;
		} else if (a == "-solid") {
			o.solid = true;
		} else if (a == "--from-table" && i + 1 < argc) {
			o.from_table = argv[++i];
		} else if (a == "--dump-mesh" && i + 1 < argc) {
			o.dump_mesh = argv[++i];
		}
	}
	if (!have_file) { printf("[Err] You didn't specify a file using -f (path). This is required. Exiting. \n"); exit(1); }
	printf("[Info] Filename: %s \n", o.filename.c_str());
	printf("[Info] Grid size: %i \n", o.gridsize);
	printf("[Info] Output format: %s \n", kFormatNames[(int)o.format]);
	printf("[Info] Using CPU-based voxelization: %s (default: No)\n", o.force_cpu ? "Yes" : "No");
	printf("[Info] Using Solid Voxelization: %s (default: No)\n", o.solid ? "Yes" : "No");
	return o;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] void die_abi(const char* where, int rc) {
	fprintf(stderr, "CUDA error at %s code=%d \"%s\" \n", where, rc, voxb200_last_error());
	exit(EXIT_FAILURE);
}
}  // namespace

int main(int argc, char* argv[]) {
	const double t_start = now_ms();
	printf("## CUDA VOXELIZER \n");
	printf("CUDA Voxelizer (B200-native build, %s), drop-in for cuda_voxelizer v0.6\n", voxb200_version());
	printf("\n## PROGRAM PARAMETERS \n");
	const Options opt = parse(argc, argv);
	fflush(stdout);

	printf("\n## READ MESH \n");
	printf("[I/O] Reading mesh from %s \n", opt.filename.c_str());
	Mesh mesh;
	std::string err;
	if (!load_mesh(opt.filename, mesh, err)) { printf("[Err] %s \n", err.c_str()); return 1; }
	printf("[Mesh] Number of triangles: %zu \n", mesh.n_faces());
	printf("[Mesh] Number of vertices: %zu \n", mesh.n_vertices());
	printf("[Mesh] Computing bbox \n");
	if (!opt.dump_mesh.empty()) {
		FILE* df = fopen(opt.dump_mesh.c_str(), "wb");
		if (!df) { printf("[Err] cannot write %s \n", opt.dump_mesh.c_str()); return 1; }
		const unsigned long long counts[2] = {mesh.n_vertices(), mesh.n_faces()};
		fwrite(counts, sizeof(counts), 1, df);
		fwrite(mesh.vertices.data(), sizeof(float), mesh.vertices.size(), df);
		fwrite(mesh.faces.data(), sizeof(int32_t), mesh.faces.size(), df);
		fclose(df);
		return 0;
	}

	printf("\n## VOXELISATION SETUP \n");
	voxb200_grid grid;
	int rc = voxb200_make_grid(mesh.bbox_min, mesh.bbox_max, opt.gridsize, mesh.n_faces(), &grid);
	if (rc) die_abi("voxb200_make_grid", rc);
	static_assert(sizeof(voxinfo) == sizeof(voxb200_grid), "voxinfo / voxb200_grid layout");      // byte-identical (static_asserts in dropin.cu)
	const voxinfo& info = *reinterpret_cast<const voxinfo*>(&grid);
	printf("[Voxelization] Bounding Box: (%f,%f,%f)-(%f,%f,%f) \n", info.bbox.min.x, info.bbox.min.y, info.bbox.min.z, info.bbox.max.x, info.bbox.max.y, info.bbox.max.z);
	printf("[Voxelization] Grid size: %i %i %i \n", info.gridsize.x, info.gridsize.y, info.gridsize.z);
	printf("[Voxelization] Triangles: %zu \n", info.n_triangles);
	printf("[Voxelization] Unit length: x: %f y: %f z: %f\n", info.unit.x, info.unit.y, info.unit.z);
	const size_t vtable_size = voxb200_table_bytes(opt.gridsize);

	if (!opt.from_table.empty()) {
		// Writer-only mode (host-side tests without a GPU): the table comes from a file instead of the voxelizer.
		std::vector<unsigned int> table(vtable_size / 4);
		FILE* tf = fopen(opt.from_table.c_str(), "rb");
		if (!tf || fread(table.data(), 1, vtable_size, tf) != vtable_size) { printf("[Err] cannot read %zu table bytes from %s \n", vtable_size, opt.from_table.c_str()); return 1; }
		fclose(tf);
		VoxelList list;
		list.gridsize = opt.gridsize;
		for (size_t w = 0; w < table.size(); w++)
			for (unsigned int bits = table[w]; bits;) {
				const int msb = 31 - __builtin_clz(bits);
				bits &= ~(1u << msb);
				list.indices.push_back((uint64_t)w * 32 + (uint64_t)(31 - msb));
			}
		printf("\n## FILE OUTPUT \n");
		switch (opt.format) {
			case Format::morton: write_binary(table.data(), vtable_size, opt.filename); break;
			case Format::binvox: write_binvox(list, info, opt.filename); break;
			case Format::obj_points: write_obj_pointcloud(list, info, opt.filename); break;
			case Format::obj_cubes: write_obj_cubes(list, info, opt.filename); break;
			case Format::vox: write_vox(list, info, opt.filename); break;
		}
		return 0;
	}
	if (opt.force_cpu) {
		printf("\n## CPU VOXELISATION \n");
		printf("[Err] -cpu: this build has no CPU voxelization path (it targets B200 GPUs only; the reference's CPU voxelizer is kept as a test oracle, not as a product path). Run without -cpu. \n");
		return 1;
	}
	printf("\n## CUDA INIT \n");
	if (!initCuda()) {
		printf("[Err] No usable CUDA GPU was found and this build has no CPU fallback. \n");
		return 1;
	}

	printf("\n## TRIANGLES TO GPU TRANSFER \n");
	const double t_up = now_ms();
	float* d_tris = nullptr;
	printf("[Mesh] Uploading %zu vertices + %zu faces; expanding to %zu bytes of triangle data on the GPU \n", mesh.n_vertices(), mesh.n_faces(), mesh.n_faces() * 36);
	rc = voxb200_upload_indexed(mesh.vertices.data(), mesh.n_vertices(), mesh.faces.data(), mesh.n_faces(), 0, &d_tris, nullptr, nullptr, nullptr);
	if (rc) die_abi("voxb200_upload_indexed", rc);
	printf("[Perf] Mesh transfer time to GPU: %.1f ms \n", now_ms() - t_up);

	printf("[Voxel Grid] Allocating %zu bytes of device memory for Voxel Grid\n", vtable_size);
	unsigned int* d_table = nullptr;
	rc = voxb200_malloc(reinterpret_cast<void**>(&d_table), vtable_size);
	if (rc) die_abi("voxb200_malloc", rc);

	printf("\n## GPU VOXELISATION \n");
	const bool morton = opt.format == Format::morton;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0, 0);
	const unsigned int flags = morton ? VOXB200_MORTON : 0u;       // clears the table itself (cudaMalloc memory is not zero)
	rc = opt.solid ? voxb200_solid(&grid, d_tris, d_table, flags, nullptr, nullptr) : voxb200_surface(&grid, d_tris, d_table, flags, nullptr, nullptr);
	if (rc) die_abi(opt.solid ? "voxb200_solid" : "voxb200_surface", rc);
	cudaEventRecord(e1, 0);
	if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "CUDA error at voxelization: %s \n", cudaGetErrorString(cudaGetLastError())); return EXIT_FAILURE; }
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, e0, e1);
	printf("[Perf] Voxelization GPU time: %.1f ms\n", ms);

	// -o morton dumps the raw table (util_io.cpp:192-200).  Every other writer only needs the SET voxels: they are
	// compacted on the GPU and only that list crosses PCIe.
	unsigned int* table_host = nullptr;
	VoxelList voxels;
	voxels.gridsize = opt.gridsize;
	const double t_rb = now_ms();
	// -o binvox on grids the device encoder covers: the run-length encoded payload is built on the GPU and is all that crosses PCIe
	std::vector<unsigned char> rle;
	const bool device_rle = opt.format == Format::binvox && opt.gridsize >= 256 && opt.gridsize % 256 == 0 && opt.gridsize <= 4096;
	if (morton) {
		// the raw table: a dense copy, or (large, mostly empty tables into aligned memory) its non-zero words expanded by host threads
		rc = voxb200_host_alloc(reinterpret_cast<void**>(&table_host), vtable_size);      // pinned and page-aligned
		if (rc) die_abi("voxb200_host_alloc", rc);
		uint64_t info[2] = {0, 0};
		rc = voxb200_download_table(d_table, vtable_size / 4, table_host, nullptr, info);
		if (rc) die_abi("voxb200_download_table", rc);
		if (info[0]) printf("[Voxel Grid] read back as %llu non-zero words \n", (unsigned long long)info[1]);
	} else if (device_rle) {
		unsigned char* d_rle = nullptr;
		size_t n_rle = 0;
		rc = voxb200_binvox_rle(d_table, opt.gridsize, &d_rle, &n_rle, nullptr);
		if (rc) die_abi("voxb200_binvox_rle", rc);
		rle.resize(n_rle);
		if (n_rle) {
			rc = voxb200_memcpy_d2h(rle.data(), d_rle, n_rle, nullptr);
			if (rc) die_abi("voxb200_memcpy_d2h", rc);
		}
		voxb200_free(d_rle);
		printf("[Voxel Grid] binvox payload: %zu bytes (run-length encoded on the GPU) \n", n_rle);
	} else {
		uint64_t* d_idx = nullptr;
		size_t n_set = 0;
		rc = voxb200_extract_voxels(d_table, vtable_size / 4, 0, &d_idx, &n_set, nullptr);
		if (rc) die_abi("voxb200_extract_voxels", rc);
		voxels.indices.resize(n_set);
		if (n_set) {
			rc = voxb200_memcpy_d2h(voxels.indices.data(), d_idx, n_set * sizeof(uint64_t), nullptr);
			if (rc) die_abi("voxb200_memcpy_d2h", rc);
		}
		voxb200_free(d_idx);
		printf("[Voxel Grid] %zu voxels set \n", n_set);
	}
	printf("[Perf] Table read-back: %.1f ms \n", now_ms() - t_rb);
	voxb200_free(d_tris);
	voxb200_free(d_table);

	printf("\n## FILE OUTPUT \n");
	const double t_out = now_ms();
	switch (opt.format) {
		case Format::morton: write_binary(table_host, vtable_size, opt.filename); break;
		case Format::binvox: if (device_rle) write_binvox_payload(rle.data(), rle.size(), info, opt.filename); else write_binvox(voxels, info, opt.filename); break;
		case Format::obj_points: write_obj_pointcloud(voxels, info, opt.filename); break;
		case Format::obj_cubes: write_obj_cubes(voxels, info, opt.filename); break;
		case Format::vox: write_vox(voxels, info, opt.filename); break;
	}
	printf("[Perf] File output: %.1f ms \n", now_ms() - t_out);
	printf("\n## STATS \n");
	printf("[Perf] Total runtime: %.1f ms \n", now_ms() - t_start);
	return 0;
}
