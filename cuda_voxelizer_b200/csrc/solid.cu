// solid.cu — solid voxelization for sm_100a (replaces voxelize_solid.cu:73-193 of the reference).
//
// The reference flips every voxel x in [0, xmax] of each accepted (y,z) column with one atomicXor
// per voxel (voxelize_solid.cu:124-137; cpu_voxelizer.cpp:297-308).  XOR is associative and
// commutative, so the same table results from
//   (1) MARK:  one atomicXor of the single bit (xmax, y, z) per accepted column sample, then
//   (2) SCAN:  a suffix-XOR along x of every (y,z) row: out(x) = XOR of marks at x' >= x,
// which turns O(G) atomics per column hit into one, and makes the fill a coalesced streaming pass.
// With the MSB-first word layout a suffix over x is a prefix over bit positions inside a word
// (w ^= w<<1; w ^= w<<2; … w<<16) plus a carry = parity of all later words of the row.
//
// Schedule: zero_kernel -> solid_tri_kernel (thread per triangle; triangles with many samples are
// queued) -> solid_coop_kernel (warp per block of samples of a queued triangle) -> solid_scan_kernel.
//
// ROW LISTS (linear order, fresh table, 128 <= G <= 4096): the marks of a row do not need the table at all until the
// fill.  Each accepted sample appends its xmax to a short per-row list (a 32-bit counter + kRowMarks 16-bit slots per
// (y,z) row, L2-resident: 20 B per row) and solid_fill_kernel builds every row straight from its list — the XOR of the
// prefix masks [0, xmax] — writing each table byte exactly once: no zero-fill, no cold-DRAM atomics, no read-back.
// A row with more than kRowMarks crossings puts the surplus marks, as single bits, into a library table that is
// all-zero between calls; the fill suffix-XORs that row in, and clears it again.  Both side buffers are left zeroed by
// the fill, so a call costs no memset.
// MARK+SCAN needs a linear table whose rows are whole words (G a power of two, 32..4096).  Morton order (whole grid)
// runs MARK+SCAN in a linear scratch table and permutes it into the morton table (linear_to_morton_kernel).  Every
// other case (odd grid sizes, morton sub-regions) takes the DIRECT mode, which flips the run itself, one atomic per
// touched word.
#include <utility>

#include "vox_internal.h"
#include "surf_micro.cuh"

namespace voxb {

constexpr int kBlock = 256;
constexpr int kSmallSide = 8;         // a thread finishes triangles whose (y,z) sample box is at most 8 x 8 itself
constexpr int kSamplesPerItem = 256;  // samples per cooperative work item (8 per lane)
constexpr int kRowMarks = 8;          // listed marks per (y,z) row (one 16-byte vector of 16-bit xmax values)

enum SolidMode { kDirect = 0, kMarkScan = 1, kRowLists = 2 };

__device__ __forceinline__ bool clip_samples_to_region(const GridParams& g, SolidSetup& s) {
	s.y0 = max(s.y0, g.ry0); s.y1 = min(s.y1, g.ry1 - 1);
	s.z0 = max(s.z0, g.rz0); s.z1 = min(s.z1, g.rz1 - 1);
	return !s.skip && s.y0 <= s.y1 && s.z0 <= s.z1;
}

// Row lists: slot `slot` of row `row` (relative to the region) takes the mark; the 9th and later marks of a row go to the overflow
// table as single bits (the region is a z-slab of whole rows: its words are (x + G * row) / 32).
__device__ __forceinline__ void row_list_commit(const GridParams& g, const RowLists& rl, unsigned int* __restrict__ overflow,
                                                unsigned int row, unsigned int slot, unsigned int xmax) {
	if (slot < (unsigned int)kRowMarks) rl.marks[(size_t)row * kRowMarks + slot] = (unsigned short)xmax;
	else {
		const unsigned long long idx = (unsigned long long)xmax + (unsigned long long)g.G * row;
		atomicXor(overflow + (idx >> 5), 1u << (31u - (unsigned int)(idx & 31ull)));
	}
}
// An ACCEPTED centre sample (y,z) of one triangle: xmax, then mark or flip.
template <int MODE, bool MORTON>
__device__ __forceinline__ void solid_emit_accepted(const SolidSetup& s, const GridParams& g, int y, int z, float py, float pz,
                                                    unsigned int* __restrict__ table, unsigned long long* __restrict__ counters, const RowLists& rl) {
	int xmax = solid_xmax(s, g, py, pz);
	// The reference leaves xmax unclamped: < 0 wraps to a 2^32-iteration out-of-bounds loop on the CPU,
	// >= G writes out of bounds.  Skip / clamp instead and count the event.
	if (xmax < 0) { atomicAdd(counters + kCtrSolidClamp, 1ull); return; }
	if (xmax > g.G - 1) { atomicAdd(counters + kCtrSolidClamp, 1ull); xmax = g.G - 1; }
	if (MODE == kRowLists) {
		// `table` is the library's overflow table here
		const unsigned int row = (unsigned int)y + (unsigned int)g.G * (unsigned int)(z - g.rz0);
		row_list_commit(g, rl, table, row, atomicAdd(rl.count + row, 1u), (unsigned int)xmax);
	} else if (MODE == kMarkScan) {
		const unsigned long long idx = voxel_index<false>(g, xmax, y, z);
		atomicXor(table + ((idx >> 5) - g.word_base), 1u << (31u - (unsigned int)(idx & 31ull)));
	} else {
		WordRun<true> run;
		const int xa = max(0, g.rx0), xb = min(xmax, g.rx1 - 1);
		for (int x = xa; x <= xb; x++) run.add(table, g, voxel_index<MORTON>(g, x, y, z));
		run.flush(table);
	}
}
// One centre sample (y,z) of one triangle: accept test, then the above.
template <int MODE, bool MORTON>
__device__ __forceinline__ void solid_emit(const SolidSetup& s, const GridParams& g, int y, int z,
                                           unsigned int* __restrict__ table, unsigned long long* __restrict__ counters, const RowLists& rl) {
	const float py = solid_center(y, g.uy), pz = solid_center(z, g.uz);
	if (!solid_sample(s, py, pz)) return;
	solid_emit_accepted<MODE, MORTON>(s, g, y, z, py, pz, table, counters, rl);
}

template <int MODE, bool MORTON, bool SOA4>
__global__ void __launch_bounds__(kBlock) solid_tri_kernel(const GridParams g, const float* __restrict__ tris,
                                                           unsigned int* __restrict__ table,
                                                           unsigned long long* __restrict__ counters,
                                                           const QueueView q, const RowLists rl) {
	grid_launch_dependents();
	// warp-private staging (72 x 16-byte cp.async per 32 triangles, no block barrier): the block-wide copy + __syncthreads it replaces
	// was 12 % of the kernel's stall samples
	__shared__ __align__(16) float stage[SOA4 ? 4 : (kBlock / 32) * 288];
	const unsigned long long i = (unsigned long long)blockIdx.x * kBlock + threadIdx.x;
	Tri t;
	const bool valid = load_tile_tri<SOA4>(g, tris, i >> 5, (int)(threadIdx.x & 31), stage + (SOA4 ? 0 : (threadIdx.x >> 5) * 288), t);
	SolidSetup s;
	bool live = false, big = false;
	unsigned int items = 0u;
	int ny = 0, nz = 0;
	if (valid) {
		shift_tri(t, g);
		solid_setup(t, g, s);
		live = clip_samples_to_region(g, s);
		if (live) {
			ny = s.y1 - s.y0 + 1; nz = s.z1 - s.z0 + 1;
			big = ny > kSmallSide || nz > kSmallSide;
			items = (unsigned int)(((long long)ny * (long long)nz + kSamplesPerItem - 1) / kSamplesPerItem);
		}
	}
	enqueue_warp(live && big, items, (unsigned int)i, q);
	if (!live || big) return;
	// Phase 1: the accept test of every sample of the (at most 8 x 8) box, outcomes kept as bit 8 (y - y0) + (z - z0): a short
	// loop body the warp runs together.  Phase 2: xmax and the list append / flips of the accepted samples only — a triangle of a
	// closed mesh owns few of its box's samples, and a warp that entered the emission once per sample ran it with a few lanes.
	unsigned long long acc = 0ull;
	for (int dy = 0; dy < ny; dy++) {
		const float py = solid_center(s.y0 + dy, g.uy);
		unsigned int row = 0u;
		for (int dz = 0; dz < nz; dz++)
			if (solid_sample(s, py, solid_center(s.z0 + dz, g.uz))) row |= 1u << dz;
		acc |= (unsigned long long)row << (8 * dy);
	}
	if (MODE == kRowLists) {
		// the append's returning atomicAdd is a round trip to L2: the slot store of one sample is issued after the NEXT sample's
		// atomic, so the trip overlaps the next xmax instead of stalling the thread
		unsigned int pend_row = 0u, pend_slot = 0u, pend_x = 0u;
		bool pend = false;
		while (acc) {
			const int k = __ffsll((long long)acc) - 1;
			acc &= acc - 1ull;
			const int y = s.y0 + (k >> 3), z = s.z0 + (k & 7);
			int xmax = solid_xmax(s, g, solid_center(y, g.uy), solid_center(z, g.uz));
			if (xmax < 0) { atomicAdd(counters + kCtrSolidClamp, 1ull); continue; }           // as solid_emit_accepted
			if (xmax > g.G - 1) { atomicAdd(counters + kCtrSolidClamp, 1ull); xmax = g.G - 1; }
			const unsigned int row = (unsigned int)y + (unsigned int)g.G * (unsigned int)(z - g.rz0);
			const unsigned int slot = atomicAdd(rl.count + row, 1u);
			if (pend) row_list_commit(g, rl, table, pend_row, pend_slot, pend_x);
			pend_row = row; pend_slot = slot; pend_x = (unsigned int)xmax; pend = true;
		}
		if (pend) row_list_commit(g, rl, table, pend_row, pend_slot, pend_x);
		return;
	}
	while (acc) {
		const int k = __ffsll((long long)acc) - 1;
		acc &= acc - 1ull;
		const int y = s.y0 + (k >> 3), z = s.z0 + (k & 7);
		solid_emit_accepted<MODE, MORTON>(s, g, y, z, solid_center(y, g.uy), solid_center(z, g.uz), table, counters, rl);
	}
}

template <int MODE, bool MORTON, bool SOA4>
__global__ void __launch_bounds__(kBlock) solid_coop_kernel(const GridParams g, const float* __restrict__ tris,
                                                            unsigned int* __restrict__ table,
                                                            unsigned long long* __restrict__ counters,
                                                            const QueueView q, const RowLists rl) {
	grid_launch_dependents();
	grid_dependency_wait();                  // the queue is the per-triangle kernel's output
	const unsigned long long packed = *q.cursor;
	const unsigned int n_entries = (unsigned int)(packed >> 32);
	const unsigned int n_items = (unsigned int)packed;
	const int lane = threadIdx.x & 31;
	const unsigned int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
	const unsigned int n_warps = (gridDim.x * kBlock) >> 5;
	for (unsigned int item = warp; item < n_items; item += n_warps) {
		const uint2 e = __ldg(&q.entries[find_slot(q, item, n_entries, n_items)]);
		Tri t;
		if (SOA4) load_tri_soa4(tris, g.n_tris, e.x, t); else load_tri_aos(tris, e.x, t);
		shift_tri(t, g);
		SolidSetup s;
		solid_setup(t, g, s);
		clip_samples_to_region(g, s);
		const int nz = s.z1 - s.z0 + 1;
		const long long samples = (long long)(s.y1 - s.y0 + 1) * (long long)nz;
		const long long k0 = (long long)(item - e.y) * kSamplesPerItem;
		const long long k1 = min(samples, k0 + (long long)kSamplesPerItem);
		for (long long k = k0 + lane; k < k1; k += 32) {
			const int y = s.y0 + (int)(k / nz), z = s.z0 + (int)(k % nz);
			solid_emit<MODE, MORTON>(s, g, y, z, table, counters, rl);
		}
	}
}

// (Both scan kernels also run in place, marks == out: no __restrict__, no read-only loads.)
// Suffix-XOR along x of every row, one lane per word: rows of `seg` = G/32 <= 32 words (a power of two).  Used for
// G = 32, 64 (rows shorter than 16 bytes) and as the fallback when a table pointer is not 16-byte aligned.
template <bool XOR_INTO>
__global__ void __launch_bounds__(kBlock) solid_scan_narrow_kernel(const unsigned int* marks, unsigned int* out, size_t n_words, int seg) {
	const size_t at = (size_t)blockIdx.x * kBlock + threadIdx.x;
	const int pos = (int)(threadIdx.x & 31) & (seg - 1);
	const unsigned int w = at < n_words ? marks[at] : 0u;
	unsigned int v = w;
	v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
	const unsigned int par = __popc(w) & 1u;
	unsigned int suf = par;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned int dn = __shfl_down_sync(0xffffffffu, suf, d);
		if (pos + d < seg) suf ^= dn;
	}
	if (suf ^ par) v = ~v;
	if (at < n_words) out[at] = XOR_INTO ? (out[at] ^ v) : v;
}

// Suffix-XOR along x of every row, rows of at least 4 words (G >= 128).  A lane owns 4 consecutive words (one
// 16-byte load and store), a row is `lanes_per_row` = G/128 <= 32 consecutive lanes of a warp (G <= 4096).
// kScanUnroll independent row groups per thread keep enough 16-byte requests in flight to stream at HBM speed.
constexpr int kScanUnroll = 4;
template <bool XOR_INTO>
__global__ void __launch_bounds__(kBlock) solid_scan_kernel(const uint4* marks, uint4* out, size_t n_vec, int lanes_per_row) {
	const int lane = threadIdx.x & 31;
	const int width = lanes_per_row;                              // lanes of this warp that share a row (<= 32)
	const int pos = lane & (width - 1);
	// a warp owns kScanUnroll consecutive 32-lane spans
	const size_t warp = ((size_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
	const size_t span0 = warp * (size_t)kScanUnroll;
	{
		uint4 m[kScanUnroll];
#pragma unroll
		for (int u = 0; u < kScanUnroll; u++) {
			const size_t at = (span0 + u) * 32 + lane;
			m[u] = at < n_vec ? marks[at] : make_uint4(0u, 0u, 0u, 0u);
		}
#pragma unroll
		for (int u = 0; u < kScanUnroll; u++) {
			const size_t at = (span0 + u) * 32 + lane;
			unsigned int w[4] = {m[u].x, m[u].y, m[u].z, m[u].w};
			const unsigned int par = (__popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3])) & 1u;
			unsigned int suf = par;                          // inclusive suffix parity over the row's lanes
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const unsigned int dn = __shfl_down_sync(0xffffffffu, suf, d);
				if (d < width && pos + d < width) suf ^= dn;
			}
			unsigned int carry = suf ^ par;                  // parity of everything after this lane's 4 words
#pragma unroll
			for (int c = 3; c >= 0; c--) {
				unsigned int v = w[c];
				v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
				if (carry) v = ~v;
				carry ^= __popc(w[c]) & 1u;
				w[c] = v;
			}
			if (at < n_vec) {
				uint4 r = make_uint4(w[0], w[1], w[2], w[3]);
				if (XOR_INTO) { const uint4 o = out[at]; r.x ^= o.x; r.y ^= o.y; r.z ^= o.z; r.w ^= o.w; }
				out[at] = r;
			}
		}
	}
}

// ROW LISTS fill: every table byte written once, from the row's list.  A lane owns 4 consecutive words (one 16-byte
// store) = 128 voxels; a row is `lanes_per_row` = G/128 <= 32 consecutive lanes of a warp.  The row is the XOR of the prefix
// masks [0, m] of its marks m, so a lane's 128 voxels are: all ones iff an odd number of marks lie at or beyond the lane's
// end, XOR the prefix masks of the marks INSIDE the lane's range.  Nearly every lane has no mark inside and few have two, so
// the kernel counts (one compare per mark), applies at most one in-range mark with three instructions per word, and leaves
// lanes with several in-range marks to a general loop that the warp only enters when one of its lanes needs it.
// Slots beyond a row's count hold 0xffff (= -1, a mark that covers nothing), so no slot needs a validity test.
// Rows that overflowed their list (count > kRowMarks) add the suffix-XOR of their row of the overflow table, which is
// cleared again on the way.  Counters and slots are reset for the next call.
constexpr int kFillUnroll = 4;
__device__ __forceinline__ int mark_at(const uint4& mk, int i) {            // slot i of a row's list, sign-extended (0xffff = -1: no mark)
	const unsigned int pair = i < 4 ? (i < 2 ? mk.x : mk.y) : (i < 6 ? mk.z : mk.w);
	return (i & 1) ? ((int)pair >> 16) : (int)(short)(pair & 0xffffu);
}
// n_vec (16-byte pieces of the region) is a multiple of 32 * kFillUnroll (G >= 128 a power of two) and fits 32 bits (G <= 4096).
__global__ void __launch_bounds__(kBlock) solid_fill_kernel(uint4* __restrict__ out, unsigned int n_vec, int row_shift, const RowLists rl,
                                                            uint4* __restrict__ overflow) {
	const int lane = threadIdx.x & 31;
	const int width = 1 << row_shift;                             // lanes per row
	const int pos = lane & (width - 1);
	const int lo = pos << 7;                                      // the lane's first x
	// a warp owns kFillUnroll consecutive 32-lane spans; all their loads are issued before any is used
	const unsigned int at0 = ((blockIdx.x * (unsigned int)kBlock + threadIdx.x) >> 5) * (unsigned int)(32 * kFillUnroll) + (unsigned int)lane;
	grid_dependency_wait();                                       // the row lists are the mark kernels' output
	if (at0 >= n_vec) return;                                     // whole warps only
	unsigned int cnt[kFillUnroll];
	uint4 mk[kFillUnroll];
	uint4* marks4 = reinterpret_cast<uint4*>(rl.marks);
#pragma unroll
	for (int u = 0; u < kFillUnroll; u++) {
		const unsigned int row = (at0 + 32u * u) >> row_shift;
		cnt[u] = rl.count[row];
		mk[u] = marks4[row];
	}
	__syncwarp();                                                 // every lane of a row has its copy before lane 0 of the row resets the list
#pragma unroll
	for (int u = 0; u < kFillUnroll; u++) {
		const unsigned int at = at0 + 32u * u;
		const unsigned int c = cnt[u];
		const unsigned int cmax = __reduce_max_sync(0xffffffffu, c);
		unsigned int w[4];
		bool fast = cmax <= 2u;                                   // warp-uniform: closed surfaces cross most rows twice
		if (fast) {
			const int d0 = mark_at(mk[u], 0) - lo, d1 = mark_at(mk[u], 1) - lo;
			const bool in0 = (unsigned int)d0 < 128u, in1 = (unsigned int)d1 < 128u;
			fast = !__any_sync(0xffffffffu, in0 && in1);
			if (fast) {
				// all ones iff an odd number of marks lie at or beyond the lane's end; the mark inside, at offset f, covers the first
				// f - 32 j + 1 bits of word j (word j holds x = lo + 32 j + (0..31), MSB first); f = -256: none
				const unsigned int fill = ((d0 >= 128) != (d1 >= 128)) ? 0xffffffffu : 0u;
				const int f = in0 ? d0 : (in1 ? d1 : -256);
#pragma unroll
				for (int j = 0; j < 4; j++) w[j] = fill ^ ~__funnelshift_rc(0xffffffffu, 0u, (unsigned int)max(f - 32 * j + 1, 0));
			}
		}
		if (!fast) {
			// the general form: every listed mark against every word (empty slots cover nothing)
#pragma unroll
			for (int j = 0; j < 4; j++) w[j] = 0u;
#pragma unroll
			for (int i = 0; i < kRowMarks; i++) {
				const int t = mark_at(mk[u], i) - lo;
#pragma unroll
				for (int j = 0; j < 4; j++) w[j] ^= ~__funnelshift_rc(0xffffffffu, 0u, (unsigned int)max(t - 32 * j + 1, 0));
			}
			if (cmax > (unsigned int)kRowMarks) {
				// surplus marks of overflowed rows: suffix-XOR of their overflow-table row (the other rows of the warp contribute zeros)
				const bool ovf = c > (unsigned int)kRowMarks;
				const uint4 o = ovf ? overflow[at] : make_uint4(0u, 0u, 0u, 0u);
				unsigned int s[4] = {o.x, o.y, o.z, o.w};
				const unsigned int par = (__popc(s[0]) + __popc(s[1]) + __popc(s[2]) + __popc(s[3])) & 1u;
				unsigned int suf = par;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const unsigned int dn = __shfl_down_sync(0xffffffffu, suf, d);
					if (d < width && pos + d < width) suf ^= dn;
				}
				unsigned int carry = suf ^ par;
#pragma unroll
				for (int j = 3; j >= 0; j--) {
					unsigned int v = s[j];
					v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
					if (carry) v = ~v;
					carry ^= __popc(s[j]) & 1u;
					w[j] ^= v;
				}
				if (ovf) overflow[at] = make_uint4(0u, 0u, 0u, 0u);
			}
		}
		out[at] = make_uint4(w[0], w[1], w[2], w[3]);
		if (pos == 0 && c != 0u) {
			rl.count[at >> row_shift] = 0u;
			marks4[at >> row_shift] = make_uint4(~0u, ~0u, ~0u, ~0u);
		}
	}
}

// Linear table -> morton table.  A morton word holds a 4(x) x 4(y) x 2(z) brick: in-word index m = x0 + 2 y0 + 4 z0 + 8 x1
// + 16 y1 (bit 31 - m).  One thread per output word gathers the brick's eight x-nibbles from the linear rows.
// Used by the solid path in morton order: the column scan needs linear rows, so the fill runs in a linear scratch
// table and is permuted once, instead of flipping O(G) morton-scattered voxels per column hit.
__device__ __forceinline__ unsigned int compact3_dev(unsigned long long m) {        // bits 0,3,6,... of m
	m &= 0x1249249249249249ull;
	m = (m | (m >> 2)) & 0x10c30c30c30c30c3ull;
	m = (m | (m >> 4)) & 0x100f00f00f00f00full;
	m = (m | (m >> 8)) & 0x001f0000ff0000ffull;
	m = (m | (m >> 16)) & 0x001f00000000ffffull;
	m = (m | (m >> 32)) & 0x00000000001fffffull;
	return (unsigned int)m;
}
template <bool XOR_INTO>
__global__ void __launch_bounds__(kBlock) linear_to_morton_kernel(const unsigned int* __restrict__ lin, unsigned int* __restrict__ out,
                                                                  size_t n_words, int G) {
	const size_t w = (size_t)blockIdx.x * kBlock + threadIdx.x;
	if (w >= n_words) return;
	const unsigned long long code = (unsigned long long)w << 5;                      // morton code of the brick's first voxel
	const unsigned int X = compact3_dev(code), Y = compact3_dev(code >> 1), Z = compact3_dev(code >> 2);
	const unsigned int Gw = (unsigned int)G >> 5;
	unsigned int word = 0u;
#pragma unroll
	for (int dz = 0; dz < 2; dz++)
#pragma unroll
		for (int dy = 0; dy < 4; dy++) {
			const size_t at = ((size_t)(Z + dz) * G + (Y + dy)) * Gw + (X >> 5);
			const unsigned int nib = (__ldg(lin + at) >> (28u - (X & 31u))) & 0xfu;      // x = X..X+3, X at the nibble's MSB
			const int c = 2 * (dy & 1) + 16 * (dy >> 1) + 4 * dz;
			word |= ((nib >> 2) << (30 - c)) | ((nib & 3u) << (22 - c));
		}
	out[w] = XOR_INTO ? (out[w] ^ word) : word;
}

// ------------------------------------------------------------------------------------------------
template <int MODE, bool MORTON, bool SOA4>
static cudaError_t run_solid_marks(Workspace& ws, const GridParams& g, const float* d_tris, unsigned int* d_marks, cudaStream_t st) {
	const unsigned int blocks = (unsigned int)((g.n_tris + kBlock - 1) / kBlock);
	RowLists rl;
	rl.count = ws.row_count; rl.marks = ws.row_marks;
	solid_tri_kernel<MODE, MORTON, SOA4><<<blocks, kBlock, 0, st>>>(g, d_tris, d_marks, ws.counters, ws.view(), rl);
	g_launch_count++;
	prof_mark(ws, 2, st);
	cudaError_t err = cudaGetLastError();
	if (err != cudaSuccess) return err;
	static int per_sm = 0;
	if (per_sm == 0) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solid_coop_kernel<MODE, MORTON, SOA4>, kBlock, 0);
	if (err != cudaSuccess) return err;
	if (per_sm < 1) per_sm = 1;
	err = launch_dependent(solid_coop_kernel<MODE, MORTON, SOA4>, (unsigned)(ws.sm_count * per_sm), kBlock, st, g, d_tris, d_marks, ws.counters, ws.view(), rl);
	g_launch_count++;
	return err != cudaSuccess ? err : cudaGetLastError();
}

template <bool XOR_INTO>
static cudaError_t run_scan(const unsigned int* marks, unsigned int* out, size_t n_words, int seg, cudaStream_t st) {
	const bool vec = seg >= 4 && (n_words & 3u) == 0 && ((reinterpret_cast<uintptr_t>(marks) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
	if (!vec) {
		if (seg > 32) return cudaErrorInvalidValue;         // G >= 2048 needs a 16-byte aligned table
		const unsigned int blocks = (unsigned int)((n_words + kBlock - 1) / kBlock);
		solid_scan_narrow_kernel<XOR_INTO><<<blocks, kBlock, 0, st>>>(marks, out, n_words, seg);
	} else {
		const size_t n_vec = n_words / 4;
		const int lanes_per_row = seg / 4;
		const size_t spans = (n_vec + 31) / 32;
		const size_t warps = (spans + (size_t)kScanUnroll - 1) / (size_t)kScanUnroll;
		const unsigned int blocks = (unsigned int)((warps + (kBlock / 32) - 1) / (kBlock / 32));
		solid_scan_kernel<XOR_INTO><<<blocks, kBlock, 0, st>>>(reinterpret_cast<const uint4*>(marks), reinterpret_cast<uint4*>(out), n_vec, lanes_per_row);
	}
	g_launch_count++;
	return cudaGetLastError();
}

cudaError_t launch_solid(Workspace& ws, const GridParams& g_in, const float* d_tris, unsigned int* d_table,
                         size_t region_words, const LaunchOpts& o, cudaStream_t st) {
	GridParams g = g_in;
	cudaError_t err = ensure_queue(ws, (size_t)g.n_tris);
	if (err != cudaSuccess) return err;
	const bool pow2 = (g.G & (g.G - 1)) == 0;
	const bool full_xy = g.rx0 == 0 && g.rx1 == g.G && g.ry0 == 0 && g.ry1 == g.G;
	const bool whole = full_xy && g.rz0 == 0 && g.rz1 == g.G;
	// MARK+SCAN needs linear rows of whole words.  Morton order gets it too (whole grid): fill a linear scratch table,
	// then permute it into the morton table.
	const bool via_linear = o.morton && whole && pow2 && g.G >= 32 && g.G <= 4096;
	const bool scan = (!o.morton && pow2 && g.G >= 32 && g.G <= 4096 && full_xy) || via_linear;
	// ROW LISTS: linear order into a fresh table whose rows are 1..32 lanes of 16 bytes
	const bool lists = scan && !via_linear && !o.accumulate && g.G >= 128 && g.G <= 4096 && (region_words & 3u) == 0 &&
	                   (reinterpret_cast<uintptr_t>(d_table) & 15u) == 0 && region_words / (size_t)(g.G / 32) <= 0xffffffffull &&
	                   ((region_words / 4) % (size_t)(32 * kFillUnroll)) == 0 && region_words / 4 <= 0xffffffffull;
	ws.last_row_lists = lists;
	if (lists) {
		const size_t n_rows = region_words / (size_t)(g.G / 32);
		err = ensure_row_lists(ws, n_rows, region_words, st);
		if (err != cudaSuccess) return err;
		prof_mark(ws, 0, st);
		err = cudaMemsetAsync(ws.counters, 0, kNumCounters * sizeof(unsigned long long), st);
		if (err != cudaSuccess) return err;
		prof_mark(ws, 1, st);
		ws.rows_dirty = true;                 // until the fill has been enqueued: it is what restores the lists' between-calls state
		if (g.n_tris != 0) {
			err = o.soa4 ? run_solid_marks<kRowLists, false, true>(ws, g, d_tris, ws.scratch, st) : run_solid_marks<kRowLists, false, false>(ws, g, d_tris, ws.scratch, st);
			if (err != cudaSuccess) return err;
		} else {
			prof_mark(ws, 2, st);
		}
		prof_mark(ws, 3, st);
		RowLists rl;
		rl.count = ws.row_count; rl.marks = ws.row_marks;
		const size_t n_vec = region_words / 4;
		int row_shift = 0;
		while ((128 << row_shift) < g.G) row_shift++;                 // lanes (of 16 bytes) per row = G / 128 = 2^row_shift
		const size_t fill_threads = (((n_vec + 31) / 32 + kFillUnroll - 1) / kFillUnroll) * 32;
		err = launch_dependent(solid_fill_kernel, (unsigned int)((fill_threads + kBlock - 1) / kBlock), kBlock, st, reinterpret_cast<uint4*>(d_table), (unsigned int)n_vec,
		                       row_shift, rl, reinterpret_cast<uint4*>(ws.scratch));
		g_launch_count++;
		if (err == cudaSuccess) err = cudaGetLastError();
		if (err != cudaSuccess) return err;
		ws.rows_dirty = false;
		prof_mark(ws, 4, st);
		if (ws.prof_on) ws.prof_calls++;
		return cudaSuccess;
	}
	unsigned int* marks = d_table;
	if (scan && (o.accumulate || via_linear)) {
		// marks must start from zero in a linear table: stage them in library scratch
		err = ensure_scratch(ws, region_words);
		if (err != cudaSuccess) return err;
		marks = ws.scratch;
		ws.scratch_zero = false;          // the row-list path re-clears it before relying on it
	}
	prof_mark(ws, 0, st);
	if (!o.accumulate || marks != d_table) err = launch_zero(ws, marks, region_words, st, true);       // also resets the counters
	else err = cudaMemsetAsync(ws.counters, 0, kNumCounters * sizeof(unsigned long long), st);
	if (err != cudaSuccess) return err;
	prof_mark(ws, 1, st);
	if (g.n_tris != 0) {
		if (scan) err = o.soa4 ? run_solid_marks<kMarkScan, false, true>(ws, g, d_tris, marks, st) : run_solid_marks<kMarkScan, false, false>(ws, g, d_tris, marks, st);
		else if (o.morton) err = o.soa4 ? run_solid_marks<kDirect, true, true>(ws, g, d_tris, d_table, st) : run_solid_marks<kDirect, true, false>(ws, g, d_tris, d_table, st);
		else err = o.soa4 ? run_solid_marks<kDirect, false, true>(ws, g, d_tris, d_table, st) : run_solid_marks<kDirect, false, false>(ws, g, d_tris, d_table, st);
		if (err != cudaSuccess) return err;
	} else {
		prof_mark(ws, 2, st);
	}
	prof_mark(ws, 3, st);
	const int seg = g.G / 32;
	if (via_linear) {
		if (g.n_tris != 0) {
			err = run_scan<false>(marks, marks, region_words, seg, st);
			if (err != cudaSuccess) return err;
		}
		const unsigned int blocks = (unsigned int)((region_words + kBlock - 1) / kBlock);
		if (o.accumulate) linear_to_morton_kernel<true><<<blocks, kBlock, 0, st>>>(marks, d_table, region_words, g.G);
		else linear_to_morton_kernel<false><<<blocks, kBlock, 0, st>>>(marks, d_table, region_words, g.G);
		g_launch_count++;
		err = cudaGetLastError();
		if (err != cudaSuccess) return err;
	} else if (scan && (g.n_tris != 0 || marks != d_table)) {
		err = (marks != d_table) ? run_scan<true>(marks, d_table, region_words, seg, st) : run_scan<false>(marks, d_table, region_words, seg, st);
		if (err != cudaSuccess) return err;
	}
	prof_mark(ws, 4, st);
	if (ws.prof_on) ws.prof_calls++;
	return cudaSuccess;
}

}  // namespace voxb
