// pow_quarter.h -- pow(x, 0.25) for x in (0.1, 1], bit for bit the result of the host libm the reference is
// linked with (glibc 2.39, x86-64 FMA variant = ARM Optimized Routines pow with GCC's contraction).
//
// It is the only transcendental on the path (ComputeScore, TranscriptDecider.hpp:394-401).  CUDA's pow() is a
// 2-ulp routine: it disagreed with the host often enough to flip a near-tie of the set cover on about one
// fuzzed workload in thirty (tools/gpu_fuzz.py); a correctly rounded fourth root still differs from glibc in
// ~1e-3 of the calls, because glibc's own result is only 0.54-ulp accurate.  So this restates glibc's
// evaluation: log(x) as a double-double from a 128-entry table + degree-7 polynomial, y * log(x), exp() from a
// 128-entry table + degree-5 polynomial, with every multiply-add fused exactly where the host build fuses it.
// tests/pow_check.cpp compares it with the host pow() on 2e8 inputs (0 differences); the constants come out of
// the image's libm.so.6 (pow_tables.h).  A host with a different libm would need its own tables / sequence.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#define PSC_TABLE __device__ const
#define PSC_POW_FN __device__ __forceinline__
#define PSC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PSC_MUL(a, b) __dmul_rn((a), (b))
#define PSC_ADD(a, b) __dadd_rn((a), (b))
#else
#include <cmath>
#define PSC_TABLE static const
#define PSC_POW_FN static inline
#define PSC_FMA(a, b, c) std::fma((a), (b), (c))
#define PSC_MUL(a, b) ((a) * (b))
#define PSC_ADD(a, b) ((a) + (b))
#endif
#include "pow_tables.h"

namespace pscdev {

PSC_POW_FN uint64_t pq_bits(double d)
{
#if defined(__CUDACC__)
	return (uint64_t)__double_as_longlong(d);
#else
	uint64_t u;
	memcpy(&u, &d, 8);
	return u;
#endif
}
PSC_POW_FN double pq_double(uint64_t u)
{
#if defined(__CUDACC__)
	return __longlong_as_double((long long)u);
#else
	double d;
	memcpy(&d, &u, 8);
	return d;
#endif
}

PSC_POW_FN double pow_quarter_libm(double x)
{
	const double y = 0.25;
	// ---- log(x) = k ln2 + log(c) + log1p(z/c - 1), as hi + tail ----
	uint64_t ix = pq_bits(x);
	uint64_t tmp = ix - 0x3fe6955500000000ull;
	int i = (int)((tmp >> (52 - 7)) % 128);
	int k = (int)((int64_t)tmp >> 52);
	uint64_t iz = ix - (tmp & (0xfffull << 52));
	double z = pq_double(iz);
	double kd = (double)k;
	double invc = k_pow_logtab[3 * i], logc = k_pow_logtab[3 * i + 1], logctail = k_pow_logtab[3 * i + 2];
	const double *A = k_pow_logpoly;
	double r = PSC_FMA(z, invc, -1.0);
	double t1 = PSC_FMA(kd, k_pow_ln2hi, logc);
	double t2 = PSC_ADD(t1, r);
	double lo1 = PSC_FMA(kd, k_pow_ln2lo, logctail);
	double lo2 = PSC_ADD(PSC_ADD(t1, -t2), r);
	double ar = PSC_MUL(A[0], r);
	double ar2 = PSC_MUL(r, ar);
	double ar3 = PSC_MUL(r, ar2);
	double hi = PSC_ADD(t2, ar2);
	double lo3 = PSC_FMA(ar, r, -ar2);
	double lo4 = PSC_ADD(PSC_ADD(t2, -hi), ar2);
	double q3 = PSC_FMA(r, A[6], A[5]);
	double q2 = PSC_FMA(ar2, q3, PSC_FMA(r, A[4], A[3]));
	double q1 = PSC_FMA(ar2, q2, PSC_FMA(r, A[2], A[1]));
	double lo = PSC_FMA(ar3, q1, PSC_ADD(PSC_ADD(PSC_ADD(lo1, lo2), lo3), lo4));
	double lg = PSC_ADD(hi, lo);
	double tail = PSC_ADD(PSC_ADD(hi, -lg), lo);
	// ---- y * log(x) as ehi + elo ----
	double ehi = PSC_MUL(y, lg);
	double elo = PSC_FMA(y, tail, PSC_FMA(y, lg, -ehi));
	// ---- exp(ehi + elo) = 2^(k/N) * exp(r), N = 128 ----
	const double invLn2N = k_pow_exp[0], shift = k_pow_exp[1], negLn2hiN = k_pow_exp[2], negLn2loN = k_pow_exp[3];
	const double C2 = k_pow_exp[4], C3 = k_pow_exp[5], C4 = k_pow_exp[6], C5 = k_pow_exp[7];
	double kd2 = PSC_FMA(invLn2N, ehi, shift);
	uint64_t ki = pq_bits(kd2);
	kd2 = PSC_ADD(kd2, -shift);
	double rr = PSC_FMA(kd2, negLn2loN, PSC_FMA(kd2, negLn2hiN, ehi));
	rr = PSC_ADD(rr, elo);
	uint64_t idx = 2 * (ki % 128);
	uint64_t top = ki << (52 - 7);
	double tl = pq_double(k_pow_exptab[idx]);
	uint64_t sbits = k_pow_exptab[idx + 1] + top;
	double r2 = PSC_MUL(rr, rr);
	double s2 = PSC_FMA(r2, PSC_FMA(rr, C3, C2), PSC_ADD(tl, rr));
	double t = PSC_FMA(PSC_MUL(r2, r2), PSC_FMA(rr, C5, C4), s2);
	double scale = pq_double(sbits);
	return PSC_FMA(scale, t, scale);
}

} // namespace pscdev
