// psc_device.cuh -- device-side building blocks shared by the stage-3 kernels.
//
// Every kernel is written against `Coop`, a cooperating group of G lanes that owns one work
// unit at a time: G = 32 (one warp per unit, the production setting) or G = 1 (one thread per
// unit, used by the tests to cross-check the warp-cooperative code).  Control flow inside a unit
// is uniform across the G lanes; data-parallel loops are strided by lane and floating-point
// accumulations are kept in the reference's sequential order (the parallel dimension is always
// the one orthogonal to the fp sum), so results are bit-identical for both widths.
//
// PSC_CPU_SHIM (tests/shim only, never part of libpsc.so) compiles the same source with g++ as
// a serial G = 1 program so the logic can be debugged in a container without a GPU.
#pragma once
#include <stdint.h>

#ifdef PSC_CPU_SHIM
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif

#define PSC_DEV __device__ __forceinline__
#define PSC_DEVN __device__ __noinline__

namespace pscdev {

#ifdef PSC_CPU_SHIM
static inline int popc32(uint32_t x) { return __builtin_popcount(x); }
static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }
static inline int ctz64(uint64_t x) { return __builtin_ctzll(x); }
static inline int ctz32(uint32_t x) { return __builtin_ctz(x); }
static inline uint64_t brev64(uint64_t x)
{
	x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
	return __builtin_bswap64(x);
}
#else
PSC_DEV int popc32(uint32_t x) { return __popc(x); }
PSC_DEV int popc64(uint64_t x) { return __popcll(x); }
PSC_DEV int ctz64(uint64_t x) { return __ffsll((long long)x) - 1; }
PSC_DEV int ctz32(uint32_t x) { return __ffs((int)x) - 1; }
PSC_DEV uint64_t brev64(uint64_t x) { return __brevll(x); }
#endif

// ---------------------------------------------------------------------------------------
// cooperative group of 1 or 32 lanes
// ---------------------------------------------------------------------------------------
struct Coop {
	int lane, n;

	PSC_DEV void sync() const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) __syncwarp();
#endif
	}
	PSC_DEV uint32_t ballot(bool p) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) return __ballot_sync(0xffffffffu, p);
#endif
		return p ? 1u : 0u;
	}
	PSC_DEV bool any(bool p) const { return ballot(p) != 0; }
	PSC_DEV int rank_in(uint32_t mask) const { return popc32(mask & ((1u << lane) - 1u)); }

	// 32-bit reductions use the warp-reduce unit (REDUX, sm_80+): one instruction instead of five shuffles
	PSC_DEV int sum(int v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) return __reduce_add_sync(0xffffffffu, v);
#endif
		return v;
	}
	PSC_DEV long long sum(long long v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1)
			for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
		return v;
	}
	PSC_DEV int max(int v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) return __reduce_max_sync(0xffffffffu, v);
#endif
		return v;
	}
	PSC_DEV int min(int v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) return __reduce_min_sync(0xffffffffu, v);
#endif
		return v;
	}
	// minimum of NON-NEGATIVE doubles: their bit patterns order like the values, so two 32-bit
	// warp reductions (high word, then low word among the lanes that hold the smallest high word) do it
	PSC_DEV double min_nonneg(double v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) {
			unsigned long long b = (unsigned long long)__double_as_longlong(v);
			unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
			unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
			unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
			return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
		}
#endif
		return v;
	}
	PSC_DEV double max(double v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1)
			for (int o = 16; o > 0; o >>= 1) {
				double t = __shfl_xor_sync(0xffffffffu, v, o);
				v = t > v ? t : v;
			}
#endif
		return v;
	}
	PSC_DEV double min(double v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1)
			for (int o = 16; o > 0; o >>= 1) {
				double t = __shfl_xor_sync(0xffffffffu, v, o);
				v = t < v ? t : v;
			}
#endif
		return v;
	}
	PSC_DEV uint64_t bor(uint64_t v) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1)
			for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
#endif
		return v;
	}
	template <class T> PSC_DEV T bcast(T v, int src) const
	{
#ifndef PSC_CPU_SHIM
		if (n > 1) return __shfl_sync(0xffffffffu, v, src);
#endif
		return v;
	}
	// smallest lane-local value of `key` wins; returns the winner's (key, payload) in all lanes.
	// Used for "first index with property" searches: key = index or INT_MAX.
	PSC_DEV int min_first(int key) const { return min(key); }
};

// ---------------------------------------------------------------------------------------
// bit vectors over subexons / pairs: W 64-bit words, bit i <-> element i (BitTable.hpp)
// ---------------------------------------------------------------------------------------
PSC_DEV bool bt_test(const uint64_t *b, int i) { return (b[i >> 6] >> (i & 63)) & 1; }
PSC_DEV void bt_set(uint64_t *b, int i) { b[i >> 6] |= (uint64_t)1 << (i & 63); }

// mask of word w keeping bits [s, e]  (BitTable::MaskRegionOutside, BitTable.hpp:182-213)
PSC_DEV uint64_t keep_mask(int w, int s, int e)
{
	int lo = w * 64, hi = lo + 63;
	if (hi < s || lo > e) return 0;
	uint64_t m = ~(uint64_t)0;
	if (s > lo) m &= ~(uint64_t)0 << (s - lo);
	if (e < hi) m &= ~(uint64_t)0 >> (hi - e);
	return m;
}

// TranscriptDecider::IsConstraintInTranscript (TranscriptDecider.cpp:181-233)
// 0 incompatible, 1 compatible, 2 head of the constraint compatible with a partial transcript.
PSC_DEV int compat(const uint64_t *tv, int tfirst, int tlast, bool tpartial, const uint64_t *cv, int cfirst,
                   int clast, int W)
{
	if (cfirst < tfirst || cfirst > tlast || !bt_test(tv, cfirst) || (!tpartial && !bt_test(tv, clast)))
		return 0;
	int s = cfirst, e = clast;
	bool retPartial = false;
	if (e > tlast) {
		if (!tpartial) return 0;
		e = tlast;
		retPartial = true;
	}
	int w0 = s >> 6, w1 = clast >> 6; // words outside [w0, w1] are zero on both sides after masking
	if (w1 >= W) w1 = W - 1;
	for (int w = w0; w <= w1; ++w) {
		uint64_t t = tv[w] & keep_mask(w, s, e);
		uint64_t c = cv[w];
		if (clast > tlast) c &= keep_mask(w, 0, e);
		if (t != c) return 0;
	}
	// constraint words below w0 are zero by construction (bits < first are never set)
	return retPartial ? 2 : 1;
}

// lexicographic order used by Constraints::CompSortConstraints (Constraints.hpp:248-264):
// at the first differing bit the vector WITHOUT the bit sorts first == ascending bit-reversed words.
// returns <0, 0, >0
PSC_DEV int cmp_bits_without_first(const uint64_t *a, const uint64_t *b, int W)
{
	for (int w = 0; w < W; ++w)
		if (a[w] != b[w]) {
			uint64_t ra = brev64(a[w]), rb = brev64(b[w]);
			return ra < rb ? -1 : 1;
		}
	return 0;
}

// ---------------------------------------------------------------------------------------
// cooperative bitonic sort of an index array (capacity must be >= next power of two of n;
// slots >= n are filled with -1 == +infinity).  less(a, b) must be a strict weak order on ids.
// ---------------------------------------------------------------------------------------
PSC_DEV int next_pow2(int n)
{
	int p = 1;
	while (p < n) p <<= 1;
	return p;
}

template <class Less> PSC_DEV void coop_sort(const Coop &co, int *idx, int n, Less less)
{
	if (n <= 1) return;
	int P = next_pow2(n);
	for (int i = n + co.lane; i < P; i += co.n) idx[i] = -1;
	co.sync();
	for (int k = 2; k <= P; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			// one compare-exchange per lane and step: pair t works on (i, i | j), i = t with a 0 inserted at bit log2(j)
			for (int t = co.lane; t < (P >> 1); t += co.n) {
				int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				int l = i | j;
				{
					int a = idx[i], b = idx[l];
					bool up = (i & k) == 0;
					// ascending pair: swap if a > b; descending pair: swap if b > a.  One comparator call either way
					// (x > y  <=>  less(y, x); -1 is +inf)
					int x = up ? a : b, y = up ? b : a;
					bool x_gt_y = (x < 0) ? (y >= 0) : (y < 0 ? false : less(y, x));
					if (x_gt_y) {
						idx[i] = b;
						idx[l] = a;
					}
				}
			}
			co.sync();
		}
	}
}

// the same network over 64-bit elements (slots >= n hold ~0 == +infinity): used with (sort prefix | index) words
template <class Less> PSC_DEV void coop_sort64(const Coop &co, uint64_t *v, int n, Less less)
{
	if (n <= 1) return;
	int P = next_pow2(n);
	for (int i = n + co.lane; i < P; i += co.n) v[i] = ~(uint64_t)0;
	co.sync();
	for (int k = 2; k <= P; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = co.lane; t < (P >> 1); t += co.n) {
				int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				int l = i | j;
				uint64_t a = v[i], b = v[l];
				bool up = (i & k) == 0;
				uint64_t x = up ? a : b, y = up ? b : a;
				bool x_gt_y = (x == ~(uint64_t)0) ? (y != ~(uint64_t)0) : (y == ~(uint64_t)0 ? false : less(y, x));
				if (x_gt_y) {
					v[i] = b;
					v[l] = a;
				}
			}
			co.sync();
		}
	}
}

#ifdef PSC_CPU_SHIM
template <class T> static inline T atomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomic_max(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomic_cas(int *p, int expect, int v) { int o = *p; if (o == expect) *p = v; return o; }
#else
template <class T> PSC_DEV T atomic_add(T *p, T v) { return atomicAdd(p, v); }
PSC_DEV int atomic_max(int *p, int v) { return atomicMax(p, v); }
PSC_DEV int atomic_cas(int *p, int expect, int v) { return atomicCAS(p, expect, v); }
#endif

} // namespace pscdev
