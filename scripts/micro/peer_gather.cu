// peer_gather.cu — the upload + share exchange of voxb200_voxelize_host_multi in isolation (bench infra, not product):
// N devices each copy 1/N of a host buffer H2D, then pull the other shares from their peers with cudaMemcpyPeerAsync.
// Reports per-phase times as the multi-device call measures them, plus a single pair's peer bandwidth.
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread -o peer_gather peer_gather.cu     Run: ./peer_gather [MB total] [devices]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
	const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 180) << 20;
	int have = 0;
	CK(cudaGetDeviceCount(&have));
	const int N = argc > 2 && atoi(argv[2]) < have ? atoi(argv[2]) : have;
	char* host = nullptr;
	CK(cudaHostAlloc(&host, bytes, cudaHostAllocPortable));
	memset(host, 1, bytes);
	std::vector<char*> d(N);
	std::vector<cudaStream_t> st(N);
	std::vector<cudaEvent_t> e0(N), e1(N), e2(N);
	int p2p_ok = 0;
	for (int k = 0; k < N; k++) {
		CK(cudaSetDevice(k));
		CK(cudaMalloc(&d[k], bytes));
		CK(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
		CK(cudaEventCreate(&e0[k])); CK(cudaEventCreate(&e1[k])); CK(cudaEventCreate(&e2[k]));
		for (int j = 0; j < N; j++) if (j != k) { int can = 0; cudaDeviceCanAccessPeer(&can, k, j); if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(j, 0); if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) p2p_ok++; cudaGetLastError(); } }
	}
	printf("%d devices, %zu MB, peer access enabled on %d of %d ordered pairs\n", N, bytes >> 20, p2p_ok, N * (N - 1));
	if (N >= 2) {      // one pair
		CK(cudaSetDevice(0));
		for (int r = 0; r < 3; r++) {
			CK(cudaStreamSynchronize(st[0]));
			const double t0 = now();
			CK(cudaMemcpyPeerAsync(d[0], 0, d[1], 1, bytes, st[0]));
			CK(cudaStreamSynchronize(st[0]));
			const double ms = now() - t0;
			if (r == 2) printf("one pair, device 1 -> 0, %zu MB: %.3f ms = %.0f GB/s\n", bytes >> 20, ms, bytes / ms / 1e6);
		}
	}
	for (int n = 2; n <= N; n *= 2) {
		const size_t share = bytes / n / 256 * 256;
		for (int r = 0; r < 3; r++) {
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaDeviceSynchronize()); }
			const double t0 = now();
			for (int k = 0; k < n; k++) {
				CK(cudaSetDevice(k));
				CK(cudaEventRecord(e0[k], st[k]));
				CK(cudaMemcpyAsync(d[k] + share * k, host + share * k, share, cudaMemcpyHostToDevice, st[k]));
				CK(cudaEventRecord(e1[k], st[k]));
			}
			for (int k = 0; k < n; k++) {
				CK(cudaSetDevice(k));
				for (int s = 1; s < n; s++) {
					const int j = (k + s) % n;
					CK(cudaStreamWaitEvent(st[k], e1[j], 0));
					CK(cudaMemcpyPeerAsync(d[k] + share * j, k, d[j] + share * j, j, share, st[k]));
				}
				CK(cudaEventRecord(e2[k], st[k]));
			}
			float h2d = 0, gather = 0;
			for (int k = 0; k < n; k++) { CK(cudaSetDevice(k)); CK(cudaStreamSynchronize(st[k])); float a, b; CK(cudaEventElapsedTime(&a, e0[k], e1[k])); CK(cudaEventElapsedTime(&b, e1[k], e2[k])); if (a > h2d) h2d = a; if (b > gather) gather = b; }
			const double wall = now() - t0;
			if (r == 2) printf("n=%d: H2D share (max) %.3f ms, peer all-gather (max) %.3f ms, wall %.3f ms\n", n, h2d, gather, wall);
		}
	}
	return 0;
}
