// multi.cu — multi-GPU voxelization behind the C ABI: ONE process, one host thread per device.
//
// The caller of the reference is a single-threaded main() that holds the indexed mesh and wants the table in host memory
// (main.cpp:203-222).  voxb200_voxelize_host_multi serves exactly that caller with N GPUs:
//   1. every device copies 1/N of the mesh BYTES (vertices and faces) from the host over its own PCIe link,
//   2. the shares are all-gathered device to device with cudaMemcpyPeerAsync (NVLink when peer access is available),
//   3. every device prepares / voxelizes the region it owns — z-slab d of N in linear order, the d-th aligned curve segment in
//      morton order (voxb200_partition); the kernels clip to the region, so the slabs are disjoint and need no reduction —
//   4. and copies its slab straight into its byte range of the ONE host table.
// PCIe carries every mesh byte once per box and every table byte once; nothing synchronises with the host between the first
// copy and the last except the two planning read-backs of the prepared mesh.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/voxb200.h"
#include "vox_internal.h"

using namespace voxb;

namespace {

constexpr int kMaxMulti = 16;

struct DevState {                       // persistent per device, grown on demand
	float* d_verts = nullptr; size_t verts_bytes = 0;
	int* d_faces = nullptr; size_t faces_bytes = 0;
	float* d_tris = nullptr; size_t tris_bytes = 0;
	unsigned int* d_table = nullptr; size_t table_bytes = 0;
	unsigned long long* d_bad = nullptr;    // faces with an out-of-range vertex index (expansion of the non-tile schedules)
	voxb200_mesh* mesh = nullptr;
	voxb200_grid mesh_grid{}; voxb200_region mesh_region{}; size_t mesh_verts = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	bool peers_enabled[64] = {};
	Readback rb;                            // slab -> host table (readback.cu)
};
DevState g_dev[64];

struct Barrier {
	std::mutex m;
	std::condition_variable cv;
	int n, waiting = 0, generation = 0;
	explicit Barrier(int n_) : n(n_) {}
	void arrive_and_wait() {
		std::unique_lock<std::mutex> lk(m);
		const int gen = generation;
		if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); }
		else cv.wait(lk, [&] { return gen != generation; });
	}
};

struct Shared {
	const voxb200_grid* grid;
	const float* host_verts; size_t n_verts;
	const int32_t* host_faces;
	unsigned int* host_table;
	unsigned int flags;
	int n;
	int devices[kMaxMulti];
	voxb200_region regions[kMaxMulti];
	size_t slab_offset[kMaxMulti], slab_bytes[kMaxMulti];
	Barrier barrier;
	std::mutex err_m;
	int rc = VOXB200_OK;
	char err[512] = "";
	bool failed = false;
	float phase_ms[kMaxMulti][6];
	explicit Shared(int n_) : n(n_), barrier(n_) {}
	void fail(int code, const char* msg) {
		std::lock_guard<std::mutex> lk(err_m);
		if (!failed) { failed = true; rc = code; strncpy(err, msg, sizeof(err) - 1); }
	}
};

template <typename T>
cudaError_t grow(T** p, size_t* have, size_t want) {
	if (want <= *have && *p) return cudaSuccess;
	if (*p) cudaFree(*p);
	*p = nullptr; *have = 0;
	cudaError_t e = cudaMalloc(p, want ? want : 16);
	if (e == cudaSuccess) *have = want;
	return e;
}

// byte range of share d of n over `bytes` bytes, cut on 16-byte boundaries
void share_range(size_t bytes, int d, int n, size_t* lo, size_t* hi) {
	const size_t units = (bytes + 15) / 16;
	*lo = (units * (size_t)d / (size_t)n) * 16;
	*hi = (units * (size_t)(d + 1) / (size_t)n) * 16;
	if (*hi > bytes) *hi = bytes;
	if (*lo > bytes) *lo = bytes;
}

#define STEP(call) do { if (!S.failed) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[400]; snprintf(b_, sizeof(b_), "device %d: %s: %s", dev, #call, cudaGetErrorString(e_)); S.fail(VOXB200_ECUDA, b_); } } } while (0)
#define STEP_RC(call) do { if (!S.failed) { int r_ = (call); if (r_ != VOXB200_OK) { char b_[512]; snprintf(b_, sizeof(b_), "device %d: %s", dev, voxb200_last_error()); S.fail(r_, b_); } } } while (0)

void worker(Shared& S, int d) {
	const int dev = S.devices[d];
	DevState& D = g_dev[dev];
	const size_t n_faces = S.grid->n_triangles;
	const size_t verts_bytes = S.n_verts * 3 * sizeof(float), faces_bytes = n_faces * 3 * sizeof(int);
	const bool morton = (S.flags & VOXB200_MORTON) != 0, solid = (S.flags & VOXB200_SOLID) != 0;
	GridParams g;
	size_t region_words = 0;
	STEP(cudaSetDevice(dev));
	STEP_RC(voxb200_init(dev));
	if (!S.failed) {
		int rc = abi_resolve_region(S.grid, S.n > 1 ? &S.regions[d] : nullptr, morton, &g, &region_words);
		if (rc) { char b[512]; snprintf(b, sizeof(b), "device %d: %s", dev, voxb200_last_error()); S.fail(rc, b); }
	}
	const bool tiles = !S.failed && mesh_tileable(g, S.flags & (VOXB200_SOLID | VOXB200_MORTON)) && n_faces > 0;
	if (!S.failed && !D.stream) {
		STEP(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
		for (auto& e : D.ev) STEP(cudaEventCreate(&e));
	}
	STEP(grow(&D.d_verts, &D.verts_bytes, verts_bytes));
	STEP(grow(&D.d_faces, &D.faces_bytes, faces_bytes + 16));
	STEP(grow(&D.d_table, &D.table_bytes, S.slab_bytes[d]));
	if (!tiles) STEP(grow(&D.d_tris, &D.tris_bytes, n_faces * 9 * sizeof(float) + 16));
	if (!S.failed) {
		for (int k = 0; k < S.n; k++) {
			const int peer = S.devices[k];
			if (peer == dev || D.peers_enabled[peer]) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, dev, peer) == cudaSuccess && can) {
				const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			}
			cudaGetLastError();
			D.peers_enabled[peer] = true;          // (tried once; without peer access cudaMemcpyPeerAsync stages through the host)
		}
	}
	S.barrier.arrive_and_wait();                 // every device's buffers exist before anybody copies into or out of them
	cudaStream_t st = D.stream;
	// 1. this device's share of the mesh bytes over its own PCIe link
	size_t vlo, vhi, flo, fhi;
	share_range(verts_bytes, d, S.n, &vlo, &vhi);
	share_range(faces_bytes, d, S.n, &flo, &fhi);
	int back_threads = readback_default_threads() / S.n;
	if (back_threads < 1) back_threads = 1;
	unsigned int* host_slab = (unsigned int*)((char*)S.host_table + S.slab_offset[d]);
	STEP(cudaEventRecord(D.ev[0], st));
	if (vhi > vlo) STEP(cudaMemcpyAsync((char*)D.d_verts + vlo, (const char*)S.host_verts + vlo, vhi - vlo, cudaMemcpyHostToDevice, st));
	if (fhi > flo) STEP(cudaMemcpyAsync((char*)D.d_faces + flo, (const char*)S.host_faces + flo, fhi - flo, cudaMemcpyHostToDevice, st));
	STEP(cudaEventRecord(D.ev[1], st));
	if (!S.failed && !solid) readback_prezero(D.rb, host_slab, S.slab_bytes[d] / sizeof(unsigned int), back_threads, st, S.n == 1);      // see voxb200_voxelize_host
	S.barrier.arrive_and_wait();                 // all share-done events are recorded
	// 2. all-gather of the shares, device to device
	for (int k = 1; k < S.n && !S.failed; k++) {
		const int j = (d + k) % S.n, peer = S.devices[j];
		size_t a, b;
		STEP(cudaStreamWaitEvent(st, g_dev[peer].ev[1], 0));
		share_range(verts_bytes, j, S.n, &a, &b);
		if (b > a) STEP(cudaMemcpyPeerAsync((char*)D.d_verts + a, dev, (const char*)g_dev[peer].d_verts + a, peer, b - a, st));
		share_range(faces_bytes, j, S.n, &a, &b);
		if (b > a) STEP(cudaMemcpyPeerAsync((char*)D.d_faces + a, dev, (const char*)g_dev[peer].d_faces + a, peer, b - a, st));
	}
	STEP(cudaEventRecord(D.ev[2], st));
	// 3. prepare + voxelize the region this device owns
	const voxb200_region* region = S.n > 1 ? &S.regions[d] : nullptr;
	if (tiles) {
		const bool same = D.mesh && D.mesh_verts == S.n_verts && memcmp(&D.mesh_grid, S.grid, sizeof(voxb200_grid)) == 0 &&
		                  memcmp(&D.mesh_region, &S.regions[d], sizeof(voxb200_region)) == 0;
		if (same) {
			STEP_RC(voxb200_mesh_update_indexed(D.mesh, D.d_verts, S.n_verts, D.d_faces, st));
		} else {
			if (D.mesh) { voxb200_mesh_destroy(D.mesh); D.mesh = nullptr; }
			STEP_RC(voxb200_mesh_create_indexed(S.grid, D.d_verts, S.n_verts, D.d_faces, 0u, region, &D.mesh, st));
			if (!S.failed) { D.mesh_grid = *S.grid; D.mesh_region = S.regions[d]; D.mesh_verts = S.n_verts; }
		}
		STEP(cudaEventRecord(D.ev[3], st));
		STEP_RC(voxb200_mesh_voxelize(D.mesh, D.d_table, 0u, st));
	} else {
		if (!S.failed && !D.d_bad) STEP(cudaMalloc(&D.d_bad, sizeof(unsigned long long)));
		STEP(cudaMemsetAsync(D.d_bad, 0, sizeof(unsigned long long), st));
		STEP(launch_expand_indexed(D.d_verts, D.d_faces, n_faces, S.n_verts, false, D.d_tris, st, D.d_bad));
		STEP(cudaEventRecord(D.ev[3], st));
		if (solid) STEP_RC(voxb200_solid(S.grid, D.d_tris, D.d_table, S.flags & VOXB200_MORTON, region, st));
		else STEP_RC(voxb200_surface(S.grid, D.d_tris, D.d_table, S.flags & VOXB200_MORTON, region, st));
	}
	STEP(cudaEventRecord(D.ev[4], st));
	// 4. the slab straight into its place in the one host table: a dense copy, or its non-zero words expanded by this device's share
	//    of the host threads (readback.cu)
	STEP(cudaEventSynchronize(D.ev[4]));
	const auto t_back = std::chrono::steady_clock::now();
	STEP_RC(readback_table(D.rb, D.d_table, S.slab_bytes[d] / sizeof(unsigned int), host_slab, st, back_threads));
	readback_cancel(D.rb);                       // (error paths) nothing writes the caller's table after this call
	const float back_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_back).count();
	STEP(cudaStreamSynchronize(st));
	if (!S.failed) {
		uint64_t c[4] = {0, 0, 0, 0};
		const int rc = tiles ? voxb200_mesh_counters(D.mesh, c) : voxb200_last_counters(c);
		if (rc == VOXB200_OK && c[1] == ~0ull) S.fail(VOXB200_EINVAL, "more than 2^32 (y,z) rows / sample blocks queued for the large-triangle path on one device: table contents undefined");
		unsigned long long bad_faces = 0ull;
		if (!tiles && D.d_bad && cudaMemcpy(&bad_faces, D.d_bad, sizeof(bad_faces), cudaMemcpyDeviceToHost) == cudaSuccess && bad_faces)
			S.fail(VOXB200_EINVAL, "faces with a vertex index out of range: table contents undefined");
	}
	for (int k = 0; k < 4; k++) {
		S.phase_ms[d][k] = 0.0f;
		if (!S.failed) cudaEventElapsedTime(&S.phase_ms[d][k], D.ev[k], D.ev[k + 1]);
	}
	S.phase_ms[d][4] = back_ms;                  // host clock: the copy engine and the host threads both work in this phase
	S.phase_ms[d][5] = 0.0f;
	if (!S.failed) { cudaEventElapsedTime(&S.phase_ms[d][5], D.ev[0], D.ev[4]); S.phase_ms[d][5] += back_ms; }
	cudaGetLastError();
	S.barrier.arrive_and_wait();                 // nobody's buffers are reused (next call) while a peer may still be reading them
}

}  // namespace

namespace voxb {
// voxb200_release: the multi-device state of one device
void multi_release_device(int dev) {
	if (dev < 0 || dev >= 64) return;
	DevState& D = g_dev[dev];
	if (D.mesh) voxb200_mesh_destroy(D.mesh);
	readback_free(D.rb);
	for (void* p : {(void*)D.d_verts, (void*)D.d_faces, (void*)D.d_tris, (void*)D.d_table, (void*)D.d_bad}) if (p) cudaFree(p);
	if (D.stream) cudaStreamDestroy(D.stream);
	for (auto& e : D.ev) if (e) cudaEventDestroy(e);
	D = DevState();
	cudaGetLastError();
}
}  // namespace voxb

extern "C" {

int voxb200_voxelize_host_multi(const voxb200_grid* grid, const float* host_verts, size_t n_verts, const int32_t* host_faces,
                                unsigned int* host_table, unsigned int flags, const int* devices, int n_devices, float timing_ms[8]) {
	if (!grid || !host_table || !host_verts || (!host_faces && grid->n_triangles)) return abi_fail(VOXB200_EINVAL, "NULL pointer");
	if (flags & ~(VOXB200_MORTON | VOXB200_SOLID)) return abi_fail(VOXB200_EINVAL, "voxb200_voxelize_host_multi takes VOXB200_MORTON and VOXB200_SOLID only");
	int have = 0;
	int rc = voxb200_device_count(&have);
	if (rc) return rc;
	if (have < 1) return abi_fail(VOXB200_ENODEVICE, "no CUDA device found");
	if (n_devices < 1 || n_devices > kMaxMulti) return abi_fail(VOXB200_EINVAL, "1..%d devices per call (got %d)", kMaxMulti, n_devices);
	Shared S(n_devices);
	S.grid = grid; S.host_verts = host_verts; S.n_verts = n_verts; S.host_faces = host_faces; S.host_table = host_table; S.flags = flags;
	const unsigned int G = grid->gridsize[0];
	size_t offset = 0;
	for (int d = 0; d < n_devices; d++) {
		const int dev = devices ? devices[d] : d;
		if (dev < 0 || dev >= have || dev >= 64) return abi_fail(VOXB200_EINVAL, "device %d out of range (have %d)", dev, have);
		for (int k = 0; k < d; k++) if (S.devices[k] == dev) return abi_fail(VOXB200_EINVAL, "device %d listed twice", dev);
		S.devices[d] = dev;
		rc = voxb200_partition(G, (flags & VOXB200_MORTON) != 0, d, n_devices, &S.regions[d], &S.slab_bytes[d]);
		if (rc) return rc;
		S.slab_offset[d] = offset;
		offset += S.slab_bytes[d];
	}
	int prev = -1;
	cudaGetDevice(&prev);
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> threads;
	for (int d = 1; d < n_devices; d++) threads.emplace_back(worker, std::ref(S), d);
	worker(S, 0);                                // the calling thread drives the first device
	for (auto& t : threads) t.join();
	const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (prev >= 0) cudaSetDevice(prev);
	if (S.failed) return abi_fail(S.rc, "%s", S.err);
	if (timing_ms) {
		for (int k = 0; k < 6; k++) {
			float mx = 0.0f;
			for (int d = 0; d < n_devices; d++) mx = S.phase_ms[d][k] > mx ? S.phase_ms[d][k] : mx;
			timing_ms[k] = mx;
		}
		timing_ms[6] = (float)wall_ms;
		timing_ms[7] = (float)n_devices;
	}
	return VOXB200_OK;
}

int voxb200_gather_slabs(unsigned int* const* d_slabs, const int* slab_devices, const size_t* slab_bytes, int n_slabs,
                         unsigned int* d_table, int table_device, void* stream) {
	if (!d_slabs || !slab_devices || !slab_bytes || !d_table || n_slabs < 1) return abi_fail(VOXB200_EINVAL, "NULL pointer / no slabs");
	size_t offset = 0;
	for (int k = 0; k < n_slabs; k++) {
		if (!d_slabs[k]) return abi_fail(VOXB200_EINVAL, "slab %d is NULL", k);
		const cudaError_t e = cudaMemcpyPeerAsync((char*)d_table + offset, table_device, d_slabs[k], slab_devices[k], slab_bytes[k], (cudaStream_t)stream);
		if (e != cudaSuccess) return abi_fail_cuda(e, "voxb200_gather_slabs: cudaMemcpyPeerAsync");
		offset += slab_bytes[k];
	}
	return VOXB200_OK;
}

int voxb200_host_alloc(void** host_ptr, size_t bytes) {
	if (!host_ptr) return abi_fail(VOXB200_EINVAL, "host_ptr is NULL");
	const cudaError_t e = cudaHostAlloc(host_ptr, bytes ? bytes : 16, cudaHostAllocPortable);
	if (e != cudaSuccess) return abi_fail_cuda(e, "cudaHostAlloc");
	return VOXB200_OK;
}

int voxb200_host_free(void* host_ptr) {
	if (host_ptr) {
		const cudaError_t e = cudaFreeHost(host_ptr);
		if (e != cudaSuccess) return abi_fail_cuda(e, "cudaFreeHost");
	}
	return VOXB200_OK;
}

}  // extern "C"
