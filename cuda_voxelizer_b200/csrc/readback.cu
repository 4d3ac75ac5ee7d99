// readback.cu — the table's way back to the host (main.cpp:203-222: the reference's writers read the table from host memory).
//
// A surface table is almost empty (config 4: 19.8 M set voxels in 2^33, 3 % of the words non-zero), yet the dense copy is 81 % of
// the end-to-end time: 1 GiB over one PCIe link = 19 ms, against 1.2 ms for everything the GPU computes.  So when few words are
// non-zero the table crosses the link as {word index, value} pairs (8 bytes per non-zero word, ascending) and host threads write
// the table: each takes 64-byte lines in order, streams zeros (non-temporal stores: no read-for-ownership of memory that is about
// to be overwritten) up to the next line that holds a pair, assembles that line and streams it (readback_host.cpp).  The pairs arrive in slices — one
// copy and one event per slice of the table — so the threads expand slice k while slice k+1 is still on the link.  Measured on the
// B200 box (scripts/micro/host_bw.cu): 8 threads stream 1 GiB of zeros in 5.8 ms (185 GB/s), the copy engine moves it in 19.1 ms.
// Dense tables (solid) and small ones take the plain copy.  Either way host_table ends up byte-identical to the device table.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/voxb200.h"
#include "vox_internal.h"

namespace voxb {

// readback_host.cpp
void readback_expand_slice(unsigned int* table, size_t w0, size_t w1, const void* pairs, size_t p0, size_t p1, bool lines_only);
void readback_expand_slice_sse2(unsigned int* table, size_t w0, size_t w1, const void* pairs, size_t p0, size_t p1, bool lines_only);

namespace {

constexpr int kMaxSlices = 256;
constexpr size_t kSparseMinBytes = 32u << 20;          // below this the dense copy is a millisecond: not worth two passes and threads

#define RB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return abi_fail_cuda(e_, #call); } while (0)

template <typename T>
cudaError_t grow_dev(T** p, size_t* have, size_t want) {
	if (want <= *have && *p) return cudaSuccess;
	if (*p) cudaFree(*p);
	*p = nullptr; *have = 0;
	cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), (want ? want : 1) * sizeof(T));
	if (e == cudaSuccess) *have = want;
	return e;
}
template <typename T>
cudaError_t grow_pinned(T** p, size_t* have, size_t want) {
	if (want <= *have && *p) return cudaSuccess;
	if (*p) cudaFreeHost(*p);
	*p = nullptr; *have = 0;
	cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(p), (want ? want : 1) * sizeof(T), cudaHostAllocPortable);
	if (e == cudaSuccess) *have = want;
	return e;
}

int dense_copy(Readback& rb, const unsigned int* d_table, size_t words, unsigned int* host_table, cudaStream_t st) {
	rb.last_mode = 0; rb.last_nonzero = 0;
	RB_CU(cudaMemcpyAsync(host_table, d_table, words * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	RB_CU(cudaStreamSynchronize(st));
	return VOXB200_OK;
}

}  // namespace

struct ReadbackHost;
static void settle_prezero(ReadbackHost& H);
static void delete_host(void* h);

std::atomic<int> g_readback_mode{[] { const char* e = getenv("VOXB200_READBACK"); return !e ? 0 : !strcmp(e, "dense") ? 1 : !strcmp(e, "sparse") ? 2 : 0; }()};

std::atomic<int> g_host_threads{0};

int readback_default_threads() {
	static int cached = 0;
	if (g_host_threads.load() > 0) return g_host_threads.load();
	if (cached) return cached;
	int n = 0;
	if (const char* e = getenv("VOXB200_HOST_THREADS")) n = atoi(e);
	if (n <= 0) {
		// streaming stores saturate the memory system with about half the hardware threads (16 vCPUs: 8 threads 185 GB/s, 16 threads 152)
		n = (int)std::thread::hardware_concurrency() / 2;
		if (n > 16) n = 16;
	}
	if (n < 1) n = 1;
	return cached = n;
}

void readback_free(Readback& rb) {
	if (rb.d_counts) cudaFree(rb.d_counts);
	if (rb.d_offsets) cudaFree(rb.d_offsets);
	if (rb.d_pairs) cudaFree(rb.d_pairs);
	if (rb.h_offsets) cudaFreeHost(rb.h_offsets);
	if (rb.h_pairs) cudaFreeHost(rb.h_pairs);
	if (rb.host) delete_host(rb.host);              // first: its workers may still be polling go_ev
	for (int k = 0; k < rb.n_ev; k++) cudaEventDestroy(rb.ev[k]);
	delete[] rb.ev;
	if (rb.go_ev) cudaEventDestroy(rb.go_ev);
	rb = Readback();
	cudaGetLastError();
}

// ---- host threads ------------------------------------------------------------------------------------------------------------
// A few persistent workers per Readback (per device): submit(fn) runs fn(worker) once on every worker, wait() returns when all are
// back.  The jobs loop over an atomic slice counter themselves.
class HostPool {
	std::vector<std::thread> threads_;
	std::mutex m_;
	std::condition_variable cv_, cv_done_;
	std::function<void(int)> job_;
	unsigned long long generation_ = 0;
	int running_ = 0;
	bool stop_ = false;
	void loop(int w, unsigned long long seen) {               // seen: the generation at the thread's creation (it must not run an older job)
		for (;;) {
			std::function<void(int)> job;
			{
				std::unique_lock<std::mutex> lk(m_);
				cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
				if (stop_) return;
				seen = generation_;
				job = job_;
			}
			job(w);
			{
				std::lock_guard<std::mutex> lk(m_);
				if (--running_ == 0) cv_done_.notify_all();
			}
		}
	}
public:
	int size() const { return (int)threads_.size(); }
	void resize(int n) {
		if (n == size()) return;
		shutdown();
		stop_ = false;
		const unsigned long long now = generation_;
		for (int w = 0; w < n; w++) threads_.emplace_back([this, w, now] { loop(w, now); });
	}
	void submit(std::function<void(int)> fn) {            // not re-entrant: wait() first
		std::lock_guard<std::mutex> lk(m_);
		job_ = std::move(fn);
		running_ = size();
		generation_++;
		cv_.notify_all();
	}
	void wait() {
		std::unique_lock<std::mutex> lk(m_);
		cv_done_.wait(lk, [&] { return running_ == 0; });
	}
	void shutdown() {
		{
			std::unique_lock<std::mutex> lk(m_);
			cv_done_.wait(lk, [&] { return running_ == 0; });
			stop_ = true;
			cv_.notify_all();
		}
		for (auto& t : threads_) t.join();
		threads_.clear();
	}
	~HostPool() { shutdown(); }
};

struct ReadbackHost {
	HostPool pool;
	// the pre-zero job of the call in flight
	bool active = false;
	unsigned int* table = nullptr;
	size_t words = 0;
	int n_slices = 0;
	std::atomic<int> next{0};
	std::atomic<bool> cancel{false};
	std::atomic<bool> go{false};             // the stream is past the upload (a polling worker saw go_ev fire, or the read-back has begun)
	std::vector<unsigned char> zeroed;       // per slice; a worker sets its slices' flags before it returns
};

static void delete_host(void* h) { ReadbackHost* H = static_cast<ReadbackHost*>(h); settle_prezero(*H); delete H; }
static ReadbackHost& host_of(Readback& rb) {
	if (!rb.host) rb.host = new ReadbackHost();
	return *static_cast<ReadbackHost*>(rb.host);
}
static int slices_for(size_t blocks) {
	static const int want = [] { const char* e = getenv("VOXB200_READBACK_SLICES"); const int v = e ? atoi(e) : 64; return v < 1 ? 1 : v > kMaxSlices ? kMaxSlices : v; }();
	return (int)(blocks < (size_t)want ? blocks : (size_t)want);
}
static bool sparse_eligible(const unsigned int* d_table, size_t words, const unsigned int* host_table) {
	const int force = g_readback_mode.load();
	const bool ok = words != 0 && (words & 15u) == 0 && words <= 0xffffffffull && (reinterpret_cast<uintptr_t>(host_table) & 63u) == 0 &&
	                (!d_table || (reinterpret_cast<uintptr_t>(d_table) & 15u) == 0);
	return ok && force != 1 && (words * sizeof(unsigned int) >= kSparseMinBytes || force == 2);
}
// Stops the pre-zero job (if any) and waits for its workers; afterwards H.zeroed says which slices are all-zero.
static void settle_prezero(ReadbackHost& H) {
	if (!H.active) return;
	H.cancel = true;
	H.pool.wait();
	H.active = false;
}

// A host entry point that expects a sparse table calls this before its first copy: the workers start streaming zeros over the host
// table while the GPU is still busy with the upload and the voxelization; the read-back then only writes the lines that hold
// non-zero words (and finishes whatever part of the zero-fill was still to do in the same pass as those lines).
// `after`: the stream the upload's host-to-device copies were enqueued on.  Most workers only start once the stream is past the
// point it is at now, because streaming stores and the copy engine's reads of host memory slow each other down (measured: a
// 3.3 ms upload took 7.1 ms next to eight zero-filling threads, and the whole call got slower); `allow_early` lets two of them
// start at once (measured: the upload does not notice two).
void readback_prezero(Readback& rb, unsigned int* host_table, size_t words, int host_threads, cudaStream_t after, bool allow_early) {
	static const bool off = getenv("VOXB200_NO_PREZERO") != nullptr;
	if (off || !sparse_eligible(nullptr, words, host_table)) return;
	ReadbackHost& H = host_of(rb);
	settle_prezero(H);
	const int T = host_threads > 0 ? host_threads : readback_default_threads();
	H.pool.resize(T);
	const size_t blocks = (words + kNzBlockWords - 1) / kNzBlockWords;
	H.table = host_table; H.words = words; H.n_slices = slices_for(blocks);
	H.zeroed.assign((size_t)H.n_slices, 0);
	H.next = 0; H.cancel = false; H.go = false; H.active = true;
	// "the stream is past the upload": an event the waiting workers poll (a host function in the stream would do, but its callback
	// thread wakes up late every now and then and everything enqueued behind it waits: measured as 30 ms outliers on 1 ms calls)
	int dev = -1;
	if (!rb.go_ev && cudaEventCreateWithFlags(&rb.go_ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); rb.go_ev = nullptr; }
	if (!rb.go_ev || cudaEventRecord(rb.go_ev, after) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); H.go = true; }
	cudaEvent_t go_ev = rb.go_ev;
	ReadbackHost* h = &H;
	// a few workers start at once (a gentle stream of stores costs the upload little), the rest when the upload is through
	static const int early_default = [] { const char* e = getenv("VOXB200_PREZERO_EARLY"); const int v = e ? atoi(e) : 2; return v < 0 ? 0 : v; }();
	const int early = allow_early ? early_default : 0;      // (several devices upload at once: every early worker is one too many, measured)
	H.pool.submit([h, blocks, early, go_ev, dev](int w) {
		if (w >= early && !h->go.load()) {
			if (dev >= 0) cudaSetDevice(dev);
			while (!h->go.load() && !h->cancel.load()) {
				if (cudaEventQuery(go_ev) != cudaErrorNotReady) { cudaGetLastError(); h->go = true; break; }
				std::this_thread::sleep_for(std::chrono::microseconds(20));
			}
		}
		for (;;) {
			if (h->cancel.load()) break;
			const int s = h->next.fetch_add(1);
			if (s >= h->n_slices) break;
			const size_t w0 = blocks * (size_t)s / (size_t)h->n_slices * kNzBlockWords;
			size_t w1 = blocks * (size_t)(s + 1) / (size_t)h->n_slices * kNzBlockWords;
			if (w1 > h->words) w1 = h->words;
			readback_expand_slice(h->table, w0, w1, nullptr, 0, 0, false);
			h->zeroed[(size_t)s] = 1;
		}
	});
}

// The non-zero words alone: ascending {word index, value} pairs in the Readback's pinned host buffer (valid until its next call).
int readback_pairs(Readback& rb, const unsigned int* d_table, size_t words, cudaStream_t st, const void** host_pairs, size_t* n_pairs) {
	if (words > 0x100000000ull) return abi_fail(VOXB200_EINVAL, "a table of more than 2^32 words has no 32-bit word indices: voxelize it in regions");
	if (reinterpret_cast<uintptr_t>(d_table) & 15u) return abi_fail(VOXB200_EINVAL, "the table must be 16-byte aligned");
	const size_t blocks = (words + kNzBlockWords - 1) / kNzBlockWords;
	if (blocks + 1 > rb.blocks_cap) {
		size_t a = 0, b = 0;
		RB_CU(grow_dev(&rb.d_counts, &a, blocks + 1));
		RB_CU(grow_dev(&rb.d_offsets, &b, blocks + 1));
		rb.blocks_cap = blocks + 1;
	}
	RB_CU(grow_pinned(&rb.h_offsets, &rb.h_blocks_cap, blocks + 1));
	RB_CU(launch_nz_count(d_table, words, rb.d_counts, rb.d_offsets, st));
	RB_CU(cudaMemcpyAsync(rb.h_offsets + blocks, rb.d_offsets + blocks, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	RB_CU(cudaStreamSynchronize(st));
	const unsigned long long nnz = rb.h_offsets[blocks];
	if (nnz > rb.pairs_cap) {
		uint2* p = reinterpret_cast<uint2*>(rb.d_pairs);
		RB_CU(grow_dev(&p, &rb.pairs_cap, (size_t)(nnz + nnz / 4 + 1024)));
		rb.d_pairs = p;
	}
	if (nnz > rb.h_pairs_cap) {
		uint2* p = reinterpret_cast<uint2*>(rb.h_pairs);
		RB_CU(grow_pinned(&p, &rb.h_pairs_cap, (size_t)(nnz + nnz / 4 + 1024)));
		rb.h_pairs = p;
	}
	if (nnz) {
		RB_CU(launch_nz_write(d_table, words, rb.d_offsets, rb.d_pairs, st));
		RB_CU(cudaMemcpyAsync(rb.h_pairs, rb.d_pairs, (size_t)nnz * sizeof(uint2), cudaMemcpyDeviceToHost, st));
	}
	RB_CU(cudaStreamSynchronize(st));
	rb.last_mode = 1; rb.last_nonzero = nnz;
	*host_pairs = rb.h_pairs;
	*n_pairs = (size_t)nnz;
	return VOXB200_OK;
}

// Stops a zero-fill that is still running ahead (error paths: nobody may write the caller's table after the call has returned).
void readback_cancel(Readback& rb) {
	if (rb.host) settle_prezero(*static_cast<ReadbackHost*>(rb.host));
}

int readback_table(Readback& rb, const unsigned int* d_table, size_t words, unsigned int* host_table, cudaStream_t st, int host_threads) {
	const size_t bytes = words * sizeof(unsigned int);
	const int force = g_readback_mode.load();
	static const bool debug = getenv("VOXB200_DEBUG_READBACK") != nullptr;
	static const bool portable = getenv("VOXB200_READBACK_SSE2") != nullptr;      // tests: the path of CPUs without AVX-512
	const auto t0 = std::chrono::steady_clock::now();
	auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	ReadbackHost& H = host_of(rb);
	const bool prezero = H.active && H.table == host_table && H.words == words;
	if (prezero) H.go = true;                             // (the stream has long passed the upload)
	if (H.active && !prezero) settle_prezero(H);          // somebody else's table: just stop it
	auto dense = [&]() -> int { settle_prezero(H); return dense_copy(rb, d_table, words, host_table, st); };
	if (!sparse_eligible(d_table, words, host_table)) return dense();
	const size_t blocks = (words + kNzBlockWords - 1) / kNzBlockWords;
	if (blocks + 1 > rb.blocks_cap) {
		size_t a = 0, b = 0;
		RB_CU(grow_dev(&rb.d_counts, &a, blocks + 1));
		RB_CU(grow_dev(&rb.d_offsets, &b, blocks + 1));
		rb.blocks_cap = blocks + 1;
	}
	RB_CU(grow_pinned(&rb.h_offsets, &rb.h_blocks_cap, blocks + 1));
	if (!rb.ev) {
		rb.ev = new cudaEvent_t[kMaxSlices];
		for (rb.n_ev = 0; rb.n_ev < kMaxSlices; rb.n_ev++) RB_CU(cudaEventCreateWithFlags(&rb.ev[rb.n_ev], cudaEventDisableTiming));
	}
	// pass 1: non-zero words per 8 KB block, their prefix, and the prefix on the host (1 MB for a 1 GiB table)
	RB_CU(launch_nz_count(d_table, words, rb.d_counts, rb.d_offsets, st));
	RB_CU(cudaMemcpyAsync(rb.h_offsets, rb.d_offsets, (blocks + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	RB_CU(cudaStreamSynchronize(st));
	const unsigned long long nnz = rb.h_offsets[blocks];
	const double t_count = since();
	// 8 bytes per pair against 4 bytes per word, plus the host's own pass over the table (~1/3 of the dense copy's time)
	if (force != 2 && nnz * 8ull > bytes / 3) return dense();
	if (nnz > rb.pairs_cap) {
		const size_t want = (size_t)(nnz + nnz / 4 + 1024);
		uint2* p = reinterpret_cast<uint2*>(rb.d_pairs);
		RB_CU(grow_dev(&p, &rb.pairs_cap, want));
		rb.d_pairs = p;
	}
	if (nnz > rb.h_pairs_cap) {
		uint2* p = reinterpret_cast<uint2*>(rb.h_pairs);
		RB_CU(grow_pinned(&p, &rb.h_pairs_cap, (size_t)(nnz + nnz / 4 + 1024)));
		rb.h_pairs = p;
	}
	// pass 2: the pairs, then slice by slice over the link
	RB_CU(launch_nz_write(d_table, words, rb.d_offsets, rb.d_pairs, st));
	const int n_slices = slices_for(blocks);
	std::vector<size_t> b0(n_slices + 1);
	for (int s = 0; s <= n_slices; s++) b0[s] = blocks * (size_t)s / (size_t)n_slices;
	const uint2* d_pairs = reinterpret_cast<const uint2*>(rb.d_pairs);
	uint2* h_pairs = reinterpret_cast<uint2*>(rb.h_pairs);
	for (int s = 0; s < n_slices; s++) {
		const size_t p0 = (size_t)rb.h_offsets[b0[s]], p1 = (size_t)rb.h_offsets[b0[s + 1]];
		if (p1 > p0) RB_CU(cudaMemcpyAsync(h_pairs + p0, d_pairs + p0, (p1 - p0) * sizeof(uint2), cudaMemcpyDeviceToHost, st));
		RB_CU(cudaEventRecord(rb.ev[s], st));
	}
	const double t_enqueued = since();
	// the zero-fill that ran ahead stops here: slices it finished only need their non-zero lines, the others the full pass
	settle_prezero(H);
	int n_zeroed = 0;
	if (prezero) for (int s = 0; s < n_slices; s++) n_zeroed += H.zeroed[(size_t)s];
	const double t_settled = since();
	int dev = -1;
	cudaGetDevice(&dev);
	std::atomic<int> next{0};
	std::atomic<int> cuda_error{0};
	const int T = host_threads > 0 ? host_threads : readback_default_threads();
	H.pool.resize(T);
	const unsigned long long* off = rb.h_offsets;
	cudaEvent_t* ev = rb.ev;
	const unsigned char* zeroed = prezero ? H.zeroed.data() : nullptr;
	H.pool.submit([&, off, ev, zeroed](int) {
		if (dev >= 0) cudaSetDevice(dev);
		for (;;) {
			const int s = next.fetch_add(1);
			if (s >= n_slices) break;
			const cudaError_t e = cudaEventSynchronize(ev[s]);
			if (e != cudaSuccess) { cuda_error = (int)e; break; }
			const size_t w0 = b0[s] * kNzBlockWords, w1 = b0[s + 1] * kNzBlockWords < words ? b0[s + 1] * kNzBlockWords : words;
			const bool lines_only = zeroed && zeroed[s];
			(portable ? readback_expand_slice_sse2 : readback_expand_slice)(host_table, w0, w1, h_pairs, (size_t)off[b0[s]], (size_t)off[b0[s + 1]], lines_only);
		}
	});
	H.pool.wait();
	const double t_expanded = since();
	RB_CU(cudaStreamSynchronize(st));
	if (debug) fprintf(stderr, "[voxb200 readback] %zu MB, %llu non-zero words, %d threads: count+sync %.3f ms, pairs enqueued %.3f, zero-fill ahead settled %.3f (%d of %d slices), expanded %.3f\n",
	                   bytes >> 20, nnz, T, t_count, t_enqueued, t_settled, n_zeroed, n_slices, t_expanded);
	if (cuda_error) return abi_fail_cuda((cudaError_t)cuda_error.load(), "readback: cudaEventSynchronize");
	rb.last_mode = 1; rb.last_nonzero = nnz;
	return VOXB200_OK;
}

}  // namespace voxb

// Diagnostics: the host-thread machinery of the read-back without a GPU (the CPU test-suite calls it): a zero-fill that runs ahead,
// cancelled, restarted with another thread count, and the two expansion passes over what it left.  0 = ok.
extern "C" int voxb200_selftest_host_pool(void) {
	using namespace voxb;
	const size_t words = (size_t)kNzBlockWords * 64 * 5;
	std::vector<unsigned int> raw(words + 16, 0xdeadbeefu);
	unsigned int* table = raw.data() + ((64 - (reinterpret_cast<uintptr_t>(raw.data()) & 63u)) & 63u) / 4;
	const int saved = g_readback_mode.load();
	g_readback_mode = 2;
	Readback rb;
	int rc = 0;
	for (int round = 0; round < 6 && rc == 0; round++) {
		const int threads = 1 + (round * 3) % 5;                  // 1, 4, 2, 5, 3, 1: the pool is rebuilt every round
		for (size_t i = 0; i < words; i++) table[i] = 0xdeadbeefu;
		ReadbackHost& H = host_of(rb);
		readback_prezero(rb, table, words, threads, nullptr, true);      // without a device the host function cannot be enqueued: starts at once
		if (round & 1) std::this_thread::sleep_for(std::chrono::milliseconds(2));
		settle_prezero(H);
		// finish by hand what readback_table would do: full pass over the slices that were not reached, nothing for the others
		const size_t blocks = words / kNzBlockWords;
		for (int s = 0; s < H.n_slices; s++) {
			const size_t w0 = blocks * (size_t)s / (size_t)H.n_slices * kNzBlockWords, w1 = blocks * (size_t)(s + 1) / (size_t)H.n_slices * kNzBlockWords;
			if (!H.zeroed[(size_t)s]) readback_expand_slice(table, w0, w1, nullptr, 0, 0, false);
		}
		for (size_t i = 0; i < words; i++) if (table[i] != 0u) { rc = 1 + round; break; }
		if (raw[0] != 0xdeadbeefu && table != raw.data()) rc = 100;
	}
	readback_free(rb);
	g_readback_mode = saved;
	return rc;
}

extern "C" int voxb200_set_host_threads(int n) {
	if (n < 0 || n > 256) return voxb::abi_fail(VOXB200_EINVAL, "host threads: 0 = default, 1..256 (got %d)", n);
	voxb::g_host_threads = n;
	return VOXB200_OK;
}

extern "C" int voxb200_set_readback_mode(int mode) {
	if (mode < 0 || mode > 2) return voxb::abi_fail(VOXB200_EINVAL, "read-back mode: 0 automatic, 1 dense, 2 sparse (got %d)", mode);
	voxb::g_readback_mode = mode;
	return VOXB200_OK;
}
