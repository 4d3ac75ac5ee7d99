// readback_host.cpp — the host threads' half of the sparse read-back (readback.cu): words [w0, w1) of the table from the ascending
// {word index, value} pairs that fall into them.  Zero lines and assembled lines are both written with non-temporal stores, a
// whole 64-byte line at a time, so the table memory is never read.  A line is assembled in REGISTERS: surface tables hold a
// non-zero word in two of every five lines, and a line built in a stack buffer (scalar stores, then vector loads of the same
// bytes) stalls on store-to-load forwarding for every one of them (measured: 20 GB/s per thread against 33 for plain zeros).
// Plain C++ (no CUDA): compiled by the host compiler alone, AVX-512 through a target attribute and a run-time check.
#include <immintrin.h>
#include <stddef.h>

namespace voxb {

struct Pair { unsigned int x, y; };

namespace {

inline __m128i put32(__m128i v, unsigned int k, unsigned int val) {          // lane k of v was zero
	__m128i x = _mm_cvtsi32_si128((int)val);
	switch (k) {
		case 0: break;
		case 1: x = _mm_slli_si128(x, 4); break;
		case 2: x = _mm_slli_si128(x, 8); break;
		default: x = _mm_slli_si128(x, 12); break;
	}
	return _mm_or_si128(v, x);
}

template <bool LINES_ONLY>
void expand_sse2(unsigned int* table, size_t w0, size_t w1, const Pair* pairs, size_t p0, size_t p1) {
	const __m128i z = _mm_setzero_si128();
	size_t line = w0 >> 4;
	const size_t end = w1 >> 4;
	size_t p = p0;
	while (line < end) {
		const size_t next = p < p1 ? (size_t)(pairs[p].x >> 4) : end;
		if (LINES_ONLY) line = next;
		for (; line < next; line++) {
			__m128i* q = reinterpret_cast<__m128i*>(table + (line << 4));
			_mm_stream_si128(q, z); _mm_stream_si128(q + 1, z); _mm_stream_si128(q + 2, z); _mm_stream_si128(q + 3, z);
		}
		if (line >= end) break;
		__m128i a = z, b = z, c = z, d = z;
		do {
			const unsigned int k = pairs[p].x & 15u, val = pairs[p].y;
			switch (k >> 2) {
				case 0: a = put32(a, k & 3u, val); break;
				case 1: b = put32(b, k & 3u, val); break;
				case 2: c = put32(c, k & 3u, val); break;
				default: d = put32(d, k & 3u, val); break;
			}
			p++;
		} while (p < p1 && (size_t)(pairs[p].x >> 4) == line);
		__m128i* q = reinterpret_cast<__m128i*>(table + (line << 4));
		_mm_stream_si128(q, a); _mm_stream_si128(q + 1, b); _mm_stream_si128(q + 2, c); _mm_stream_si128(q + 3, d);
		line++;
	}
	_mm_sfence();
}

template <bool LINES_ONLY>
__attribute__((target("avx512f"))) void expand_avx512(unsigned int* table, size_t w0, size_t w1, const Pair* pairs, size_t p0, size_t p1) {
	const __m512i z = _mm512_setzero_si512();
	size_t line = w0 >> 4;
	const size_t end = w1 >> 4;
	size_t p = p0;
	while (line < end) {
		const size_t next = p < p1 ? (size_t)(pairs[p].x >> 4) : end;
		if (LINES_ONLY) line = next;
		for (; line < next; line++) _mm512_stream_si512(reinterpret_cast<__m512i*>(table + (line << 4)), z);
		if (line >= end) break;
		__m512i v = z;
		do {
			v = _mm512_mask_set1_epi32(v, (__mmask16)(1u << (pairs[p].x & 15u)), (int)pairs[p].y);
			p++;
		} while (p < p1 && (size_t)(pairs[p].x >> 4) == line);
		_mm512_stream_si512(reinterpret_cast<__m512i*>(table + (line << 4)), v);
		line++;
	}
	_mm_sfence();
}

}  // namespace

// table: 64-byte aligned; w0, w1: multiples of 16 words; pairs[p0..p1): exactly the pairs with w0 <= x < w1, ascending.
// lines_only: the words are already zero — write just the lines that hold a pair.
void readback_expand_slice(unsigned int* table, size_t w0, size_t w1, const void* pairs, size_t p0, size_t p1, bool lines_only) {
	static const bool avx512 = __builtin_cpu_supports("avx512f");
	const Pair* pp = static_cast<const Pair*>(pairs);
	if (avx512) { if (lines_only) expand_avx512<true>(table, w0, w1, pp, p0, p1); else expand_avx512<false>(table, w0, w1, pp, p0, p1); }
	else { if (lines_only) expand_sse2<true>(table, w0, w1, pp, p0, p1); else expand_sse2<false>(table, w0, w1, pp, p0, p1); }
}
// (tests) the portable path, whatever the CPU
void readback_expand_slice_sse2(unsigned int* table, size_t w0, size_t w1, const void* pairs, size_t p0, size_t p1, bool lines_only) {
	const Pair* pp = static_cast<const Pair*>(pairs);
	if (lines_only) expand_sse2<true>(table, w0, w1, pp, p0, p1); else expand_sse2<false>(table, w0, w1, pp, p0, p1);
}

}  // namespace voxb
