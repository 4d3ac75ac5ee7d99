// host_bw.cu — what the HOST side of the end-to-end path can do on this box (bench infra, not product):
//   (1) zero-filling a pinned table with T threads (memset / non-temporal stores),
//   (2) D2H of the same bytes by the copy engine, alone and while the threads are zero-filling another buffer,
//   (3) N devices copying D2H at once into one pinned buffer (is the box's host memory or the links the limit?).
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread -o host_bw host_bw.cu      Run: ./host_bw [MiB] [max devices]
#include <cuda_runtime.h>
#include <immintrin.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void zero_nt(char* p, size_t n) {
	__m128i z = _mm_setzero_si128();
	for (size_t i = 0; i < n; i += 64) {
		_mm_stream_si128((__m128i*)(p + i), z); _mm_stream_si128((__m128i*)(p + i + 16), z);
		_mm_stream_si128((__m128i*)(p + i + 32), z); _mm_stream_si128((__m128i*)(p + i + 48), z);
	}
	_mm_sfence();
}

static double zero_threads(char* buf, size_t bytes, int T, bool nt) {
	std::vector<std::thread> th;
	const double t0 = now();
	for (int t = 0; t < T; t++) {
		const size_t a = (bytes / T / 4096) * 4096 * t, b = t == T - 1 ? bytes : (bytes / T / 4096) * 4096 * (t + 1);
		th.emplace_back([=] { if (nt) zero_nt(buf + a, b - a); else memset(buf + a, 0, b - a); });
	}
	for (auto& x : th) x.join();
	return now() - t0;
}

int main(int argc, char** argv) {
	const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 1024) << 20;
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	if (argc > 2 && atoi(argv[2]) < ndev) ndev = atoi(argv[2]);
	const int hw = (int)std::thread::hardware_concurrency();
	printf("host threads %d, devices %d, buffer %zu MiB\n", hw, ndev, bytes >> 20);
	char *a = nullptr, *b = nullptr;
	CK(cudaHostAlloc(&a, bytes, cudaHostAllocPortable));
	CK(cudaHostAlloc(&b, bytes, cudaHostAllocPortable));
	memset(a, 1, bytes); memset(b, 1, bytes);
	for (int nt = 0; nt < 2; nt++)
		for (int T = 1; T <= hw && T <= 64; T *= 2) {
			double best = 1e9;
			for (int r = 0; r < 3; r++) { const double s = zero_threads(a, bytes, T, nt); if (s < best) best = s; }
			printf("zero-fill %-7s T=%2d: %7.2f ms  %6.1f GB/s\n", nt ? "stream" : "memset", T, best * 1e3, bytes / best / 1e9);
		}
	std::vector<char*> d(ndev);
	std::vector<cudaStream_t> st(ndev);
	for (int k = 0; k < ndev; k++) { CK(cudaSetDevice(k)); CK(cudaMalloc(&d[k], bytes)); CK(cudaMemset(d[k], 0, bytes)); CK(cudaStreamCreate(&st[k])); }
	for (int n = 1; n <= ndev; n *= 2) {
		// n devices, each 1/n of the buffer
		double best = 1e9;
		for (int r = 0; r < 3; r++) {
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaDeviceSynchronize()); }
			const double t0 = now();
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaMemcpyAsync(a + bytes / n * k, d[k], bytes / n, cudaMemcpyDeviceToHost, st[k])); }
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaStreamSynchronize(st[k])); }
			const double s = now() - t0;
			if (s < best) best = s;
		}
		printf("D2H %zu MiB over %d device(s): %7.2f ms  %6.1f GB/s total\n", bytes >> 20, n, best * 1e3, bytes / best / 1e9);
		best = 1e9;
		for (int r = 0; r < 3; r++) {
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaDeviceSynchronize()); }
			const double t0 = now();
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaMemcpyAsync(d[k], a + bytes / n * k, bytes / n, cudaMemcpyHostToDevice, st[k])); }
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaStreamSynchronize(st[k])); }
			const double s = now() - t0;
			if (s < best) best = s;
		}
		printf("H2D %zu MiB over %d device(s): %7.2f ms  %6.1f GB/s total\n", bytes >> 20, n, best * 1e3, bytes / best / 1e9);
	}
	// D2H on device 0 while T threads zero-fill the other buffer
	CK(cudaSetDevice(0));
	for (int T : {4, 8, 16, 32}) {
		if (T > hw) break;
		CK(cudaDeviceSynchronize());
		const double t0 = now();
		CK(cudaMemcpyAsync(a, d[0], bytes, cudaMemcpyDeviceToHost, st[0]));
		const double z = zero_threads(b, bytes, T, true);
		CK(cudaStreamSynchronize(st[0]));
		const double s = now() - t0;
		printf("D2H + zero-fill(stream, T=%2d) together: zero %7.2f ms, both done %7.2f ms\n", T, z * 1e3, s * 1e3);
	}
	// scatter of a sparse word list into the zeroed table (every 32nd word), T threads
	{
		const size_t n = bytes / 4 / 32;
		std::vector<unsigned int> idx(n), val(n, 0x80000001u);
		for (size_t i = 0; i < n; i++) idx[i] = (unsigned int)(i * 32 + (i * 7) % 32);
		for (int T : {1, 4, 8, 16}) {
			if (T > hw) break;
			std::vector<std::thread> th;
			const double t0 = now();
			for (int t = 0; t < T; t++) th.emplace_back([&, t] { unsigned int* tab = (unsigned int*)a; for (size_t i = n * t / T; i < n * (t + 1) / T; i++) tab[idx[i]] = val[i]; });
			for (auto& x : th) x.join();
			printf("scatter %zu words (1 in 32) T=%2d: %7.2f ms\n", n, T, (now() - t0) * 1e3);
		}
	}
	return 0;
}
