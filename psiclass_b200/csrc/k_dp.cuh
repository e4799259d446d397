// k_dp.cuh -- candidate generation by dynamic programming for complex components:
// PickTranscriptsByDP / SolveSubTranscript / SearchSubTranscript / AugmentTranscripts /
// CoalesceSameTranscripts (TranscriptDecider.cpp:303-328, 374-1365) and the per-sample driver
// loop of Solve (TranscriptDecider.cpp:3139-3265).
//
// The two mutually recursive functions of the reference become one explicit-stack machine whose
// control state is uniform across the cooperative group; the scans over the constraint list
// (cover counting, extension collection) are strided across the lanes.  The depth-first order and
// the strict '>' updates of the reference are preserved, so the chosen paths are identical.
#pragma once

namespace pscdev {

struct DpMemo {
	double cover, minAbundance;
	// smallest normAbund among the constraints compatible with the stored path (those the recursion
	// examined): the entry stays exact for every higher threshold below it, see dp_memo_hit
	double beta;
	int first, last, cnt, timeStamp, strand;
	int iter, mode; // outer iteration / forAbundance flag the entry was computed under
};

struct DpScratch {
	// memo: f1 [n], f2 [n*n], hash [hashMax]; an entry is valid only when its epoch (kept in an array of
	// its own, zeroed by the unit before use) matches
	DpMemo *ent;
	int *eepoch;
	uint64_t *ebits;
	long long f2Base, hashBase;
	int hashMax;
	// constraints handed to the DP for the current sample (possibly truncated / down-sampled)
	uint64_t *tcb;
	int *tcf, *tcl, *tcsup;
	double *tcnorm, *tcab;
	int tcCnt;
	uint8_t *uncov; // per next-edge of the component
	int *P;         // the path stack: solve keys and search extensions share it
	int *fr;        // frames
	double *fdcover, *fdbeta;
	uint64_t *fdkey; // [(n+3)*W] key bits per owner
	uint64_t *pbits; // [(n+4)*W] owner key | walk, per path position
	int *fdmeta; // first,last,strand,#uncovered edges inside the key: per solve frame
	int *tcEnd;  // [n+1] number of constraints whose .first <= q
	uint64_t *fdbits;
	int *ext;
	int extCap;
	double *ckey;
	DpMemo *cent;
	uint64_t *cbits;
	uint8_t *cvalid;
	int cCap;
	uint64_t *txb;
	int *txf, *txl;
	int txCap;
	uint64_t *work; // [8*W]
	int *covered;
	int *sortidx;
	// second candidate buffer for the permutation step of CoalesceSameTranscripts
	uint64_t *c2b;
	int *c2f, *c2l;
	// shared-memory sort buffer of the merge kernel (null: sort indices in global memory)
	uint64_t *sort64;
	int sort64Cap;
};

// frame memory of the walk: 5 ints per path position, 10 per owner (k_dp_walk.cuh)
#define PSC_FR_INTS_TOTAL(n) (15 * (n) + 64)
__host__ __device__ inline size_t dp_scratch_layout(DpScratch *o, char *base, int n, int E, int Cmax, long long candCap)
{
	Carver cv(base);
	int W = (n + 63) >> 6;
	DpScratch s;
	// The reference uses 100003 / 1000003 slots (Q3).  A collision only evicts an entry, which is then
	// recomputed to the same value, so the table size changes no result (measured); a smaller
	// direct-mapped table keeps the per-unit memo in L2.
	s.hashMax = 8191;
	long long nEnt = (long long)n + (long long)n * n + s.hashMax;
	s.f2Base = n;
	s.hashBase = (long long)n + (long long)n * n;
	s.eepoch = cv.take<int>((size_t)nEnt);
	s.ent = cv.take<DpMemo>((size_t)nEnt);
	s.ebits = cv.take<uint64_t>((size_t)nEnt * W);
	s.tcb = cv.take<uint64_t>((size_t)Cmax * W + 1);
	s.tcf = cv.take<int>((size_t)Cmax + 1);
	s.tcl = cv.take<int>((size_t)Cmax + 1);
	s.tcsup = cv.take<int>((size_t)Cmax + 1);
	s.tcnorm = cv.take<double>((size_t)Cmax + 1);
	s.tcab = cv.take<double>((size_t)Cmax + 1);
	s.tcCnt = 0;
	s.uncov = cv.take<uint8_t>((size_t)E + 1);
	s.P = cv.take<int>((size_t)n + 4);
	s.fr = cv.take<int>((size_t)PSC_FR_INTS_TOTAL(n));
	s.fdcover = cv.take<double>((size_t)n + 3);
	s.fdbeta = cv.take<double>((size_t)n + 3);
	s.fdmeta = cv.take<int>((size_t)(n + 3) * 4);
	s.tcEnd = cv.take<int>((size_t)n + 2);
	s.fdbits = cv.take<uint64_t>((size_t)(n + 3) * W);
	s.fdkey = cv.take<uint64_t>((size_t)(n + 3) * W);
	s.pbits = cv.take<uint64_t>((size_t)(n + 4) * W);
	long long ec = (long long)(n + 2) * (Cmax + 1);
	if (ec > (long long)4 << 20) ec = (long long)4 << 20;
	s.extCap = (int)ec;
	s.ext = cv.take<int>((size_t)s.extCap);
	s.cCap = Cmax + 2;
	s.ckey = cv.take<double>((size_t)s.cCap);
	s.cent = cv.take<DpMemo>((size_t)s.cCap);
	s.cbits = cv.take<uint64_t>((size_t)s.cCap * W);
	s.cvalid = cv.take<uint8_t>((size_t)s.cCap);
	s.txCap = Cmax + 2;
	s.txb = cv.take<uint64_t>((size_t)s.txCap * W);
	s.txf = cv.take<int>((size_t)s.txCap);
	s.txl = cv.take<int>((size_t)s.txCap);
	s.work = cv.take<uint64_t>((size_t)8 * W);
	s.covered = cv.take<int>((size_t)Cmax + 1);
	long long p2 = 1;
	while (p2 < candCap + 1) p2 <<= 1;
	s.sortidx = cv.take<int>((size_t)p2 + 1);
	s.c2b = cv.take<uint64_t>((size_t)candCap * W + 1);
	s.c2f = cv.take<int>((size_t)candCap + 1);
	s.c2l = cv.take<int>((size_t)candCap + 1);
	s.sort64 = nullptr;
	s.sort64Cap = 0;
	if (o) *o = s;
	return (cv.off + 15) & ~(size_t)15;
}
// number of memo entries (the unit zeroes their epochs before the first use)
__host__ __device__ inline long long dp_memo_entries(int n) { return (long long)n + (long long)n * n + 8191; }

struct DpState {
	Coop co;
	CompView v;
	DpScratch d;
	const uint8_t *cs, *ce;
	int W;
	bool forAbundance;
	double minAbundance;
	int timeStamp, epoch, iter;
	int *err;
	struct Dp2 *z; // static-program engine (k_dp2.cuh), null for the walk engine
	bool prof; // PSC_DEBUG_DP: per-mode cycle counters
	mutable long long nStates, nLevels, nIters, nFrames, nRoots, cycRoot, cycAssign; // instrumentation (PSC_DEBUG_DP)
};

PSC_DEV bool dp_edge_uncovered(const DpState &st, int a, int b)
{
	for (int e = st.v.noff[a]; e < st.v.noff[a + 1]; ++e)
		if (st.v.nidx[e] == b) return st.d.uncov[e - st.v.noff[0]] != 0;
	return false;
}

// first index >= from whose constraint starts after `bound` (constraints are sorted by .first)
PSC_DEV int dp_tc_end(const DpState &st, int from, int bound)
{
	int e = bound >= st.v.n - 1 ? st.d.tcCnt : (bound < 0 ? 0 : st.d.tcEnd[bound]);
	return e > from ? e : from;
}

// sub = [base |] bits(p1) | bits(p2).  For components of <= 128 subexons every lane keeps its own copy
// in registers (no shared buffer, no synchronisation); wider ones go through the shared buffer.
PSC_DEV void dp_build_sub(const DpState &st, uint64_t *sub, const uint64_t *base, const int *p1, int n1, const int *p2, int n2)
{
	if (st.W <= 2) {
		uint64_t a = base ? base[0] : 0, b = (base && st.W == 2) ? base[1] : 0;
		for (int i = 0; i < n1; ++i) {
			int x = p1[i];
			if (x < 64) a |= (uint64_t)1 << x;
			else b |= (uint64_t)1 << (x - 64);
		}
		for (int i = 0; i < n2; ++i) {
			int x = p2[i];
			if (x < 64) a |= (uint64_t)1 << x;
			else b |= (uint64_t)1 << (x - 64);
		}
		sub[0] = a;
		if (st.W == 2) sub[1] = b;
		return;
	}
	st.co.sync();
	if (st.co.lane == 0) {
		for (int w = 0; w < st.W; ++w) sub[w] = base ? base[w] : 0;
		for (int i = 0; i < n1; ++i) bt_set(sub, p1[i]);
		for (int i = 0; i < n2; ++i) bt_set(sub, p2[i]);
	}
	st.co.sync();
}

// the three cover scans of the reference (.cpp:463-490, 569-596, 732-755), lane-parallel.
// All reductions are order-free: a count of integers, a minimum, or "any".
PSC_DEV double dp_scan_cover(const DpState &st, const uint64_t *tb, int tf, int tl, int tcStart, int bound, double cover,
                             bool &belowMin, double &beta)
{
	int end = dp_tc_end(st, tcStart, bound);
	if (end <= tcStart) return cover; // nothing starts inside the range: no lane has work, no reduction
	bool below = false;
	int cnt = 0, ccnt = 0;
	double mn = 1e300, mnAll = 1e300;
	for (int i = tcStart + st.co.lane; i < end; i += st.co.n) {
		if (compat(tb, tf, tl, false, st.d.tcb + (long long)i * st.W, st.d.tcf[i], st.d.tcl[i], st.W) != 1) continue;
		if (st.d.tcnorm[i] <= st.minAbundance) { below = true; continue; }
		++ccnt;
		if (st.d.tcnorm[i] < mnAll) mnAll = st.d.tcnorm[i];
		if (st.d.tcab[i] <= 0) continue;
		++cnt;
		if (st.d.tcnorm[i] < mn) mn = st.d.tcnorm[i];
	}
	// the reference breaks at the FIRST compatible constraint with normAbund <= minAbundance, whatever
	// was counted before, and returns -2: "any" is enough
	if (st.co.any(below)) {
		belowMin = true;
		return -2;
	}
	// beta: smallest normAbund among the compatible constraints, depleted ones included
	if (!st.forAbundance && st.co.sum(ccnt) > 0) {
		mnAll = st.co.min_nonneg(mnAll);
		if (mnAll < beta) beta = mnAll;
	}
	cnt = st.co.sum(cnt);
	if (st.forAbundance) {
		if (cnt > 0) {
			mn = st.co.min_nonneg(mn);
			if (cover == 0 || mn < cover) cover = mn;
		}
		return cover;
	}
	return cover + cnt;
}

// ---- components of <= 256 subexons: bit vectors of W <= NW words held in registers (NW = 2 or 4: few
// instantiations, because warps of one SM running different copies of the walk thrash the instruction cache) ----
// (every loop below has a compile-time trip count and is unrolled, so w[] never becomes an array in memory)
template <int NW> struct BitsR {
	uint64_t w[NW > 0 ? NW : 1];
	PSC_DEV void clear()
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] = 0;
	}
	// W <= NW words in memory; the words above W stay zero
	PSC_DEV void load(const uint64_t *p, int W)
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] = (i == 0 || i < W) ? p[i < W ? i : 0] : 0;
	}
	PSC_DEV void store(uint64_t *p, int W) const
	{
#pragma unroll
		for (int i = 0; i < NW; ++i)
			if (i == 0 || i < W) p[i] = w[i];
	}
	PSC_DEV void set(int q)
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] |= (q >> 6) == i ? (uint64_t)1 << (q & 63) : 0;
	}
	PSC_DEV bool test(int q) const
	{
		uint64_t x = w[NW > 0 ? NW - 1 : 0];
#pragma unroll
		for (int i = 0; i + 1 < NW; ++i)
			if ((q >> 6) == i) x = w[i];
		return (x >> (q & 63)) & 1;
	}
	PSC_DEV void or_with(const BitsR &o)
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] |= o.w[i];
	}
	PSC_DEV void and_not(const BitsR &a, const BitsR &b)
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] = a.w[i] & ~b.w[i];
	}
	// mask of the bits [s, e], 0 <= s, e < 64 * NW (empty when s > e)
	PSC_DEV void range(int s, int e)
	{
#pragma unroll
		for (int i = 0; i < NW; ++i) {
			const int lo = 64 * i, hi = lo + 63;
			// bits >= s: all of the word when s <= lo (the shift does that by itself in word 0), none when s > hi
			uint64_t a = ~(uint64_t)0 << ((s - lo) & 63);
			if (i > 0 && s <= lo) a = ~(uint64_t)0;
			if (i + 1 < NW && s > hi) a = 0;
			// bits <= e: all of the word when e >= hi (the shift does that in the last word), none when e < lo
			uint64_t b = ~(uint64_t)0 >> ((hi - e) & 63);
			if (i + 1 < NW && e >= hi) b = ~(uint64_t)0;
			if (i > 0 && e < lo) b = 0;
			w[i] = a & b;
		}
	}
	// (*this & m) == x
	PSC_DEV bool masked_equals(const BitsR &m, const BitsR &x) const
	{
		bool eq = true;
#pragma unroll
		for (int i = 0; i < NW; ++i) eq = eq && (w[i] & m.w[i]) == x.w[i];
		return eq;
	}
};

// compat() without loops or early exits: same decisions, see the case analysis there
template <int NW> PSC_DEV int compat_r(const BitsR<NW> &t, int tf, int tl, bool tpartial, const BitsR<NW> &c, int cf, int cl)
{
	bool part = cl > tl;
	int e = part ? tl : cl;
	BitsR<NW> m;
	m.range(cf, e);
	bool ok = cf >= tf && cf <= tl && t.test(cf) && (tpartial || (!part && t.test(cl)));
	// a constraint has no bits below its first subexon, so keep(0, e) and keep(cf, e) select the same bits of it
	bool eq = true;
#pragma unroll
	for (int i = 0; i < NW; ++i) {
		uint64_t x = part ? (c.w[i] & m.w[i]) : c.w[i];
		eq = eq && (t.w[i] & m.w[i]) == x;
	}
	return ok && eq ? (part ? 2 : 1) : 0;
}

// dp_scan_cover with the path bits in registers; the per-constraint decisions are selects, not
// branches.  The warp reductions are kept to one vote, one packed add and -- only when some lane can
// lower the running minimum -- one minimum.  Deliberately NOT unrolled: the walk is bound by
// instruction fetch as much as by latency, and every variant with more code per scan (two constraints
// per lane, a special case for short ranges) measured slower on B200 (total DP cycles of the bench
// workload: 1.79e12 with both, 1.57e12 without the special case, 1.24e12 with the plain loop).
#ifndef PSC_SCANR_INLINE
#define PSC_SCANR_INLINE PSC_DEV
#endif
template <int NW>
PSC_SCANR_INLINE double dp_scan_cover_r(const DpState &st, const BitsR<NW> &t, int tf, int tl, int tcStart, int bound, double cover,
                               bool &belowMin, double &beta)
{
	int end = dp_tc_end(st, tcStart, bound);
	if (end <= tcStart) return cover;
	const DpScratch &d = st.d;
	const int L = st.co.n, W = st.W;
	const double minAb = st.minAbundance;
	bool below = false;
	int cnt = 0, ccnt = 0;
	double mn = 1e300, mnAll = 1e300;
	{
		for (int i0 = tcStart + st.co.lane; i0 < end; i0 += L) {
			BitsR<NW> y0;
			y0.load(d.tcb + (long long)i0 * W, W);
			double n0 = d.tcnorm[i0], x0 = d.tcab[i0];
			bool c0 = compat_r<NW>(t, tf, tl, false, y0, d.tcf[i0], d.tcl[i0]) == 1;
			bool lo0 = n0 <= minAb;
			below = below || (c0 && lo0);
			bool in0 = c0 && !lo0;
			ccnt += in0 ? 1 : 0;
			if (in0 && n0 < mnAll) mnAll = n0;
			bool p0 = in0 && x0 > 0;
			cnt += p0 ? 1 : 0;
			if (p0 && n0 < mn) mn = n0;
		}
	}
	// the reference breaks at the FIRST compatible constraint with normAbund <= minAbundance, whatever
	// was counted before, and returns -2: "any" is enough
	if (st.co.any(below)) {
		belowMin = true;
		return -2;
	}
	// both counts in one reduction when they cannot overflow their 15-bit fields
	if (d.tcCnt < 32768) {
		int both = st.co.sum((ccnt << 15) | cnt);
		ccnt = both >> 15;
		cnt = both & 32767;
	} else {
		ccnt = st.co.sum(ccnt);
		cnt = st.co.sum(cnt);
	}
	// beta: smallest normAbund among the compatible constraints, depleted ones included
	if (!st.forAbundance && ccnt > 0 && st.co.any(mnAll < beta)) {
		mnAll = st.co.min_nonneg(mnAll);
		if (mnAll < beta) beta = mnAll;
	}
	if (st.forAbundance) {
		if (cnt > 0) {
			mn = st.co.min_nonneg(mn);
			if (cover == 0 || mn < cover) cover = mn;
		}
		return cover;
	}
	return cover + cnt;
}

PSC_DEV void dp_bits_from_path(const DpState &st, uint64_t *dst, const int *p, int cnt, bool clear)
{
	// uniform scalar code executed by lane 0 only; callers sync
	if (st.co.lane != 0) return;
	if (clear)
		for (int w = 0; w < st.W; ++w) dst[w] = 0;
	for (int i = 0; i < cnt; ++i) bt_set(dst, p[i]);
}

#define PSC_DP_HASH 8191u
PSC_DEV long long dp_memo_index(const DpState &st, const int *visit, int vcnt)
{
	if (vcnt == 1) return visit[0];
	if (vcnt == 2) return st.d.f2Base + (long long)visit[0] * st.v.n + visit[1]; // seCnt <= 10000 is enforced by the host
	// key < 8191 and seCnt <= 10000: the products fit 32 bits, and the modulus is a compile-time constant
	uint32_t key = 0, n = (uint32_t)st.v.n;
	for (int i = 0; i < vcnt; ++i) key = (key * n + (uint32_t)visit[i]) % PSC_DP_HASH;
	return st.d.hashBase + key;
}

// Memo validity.  The reference accepts an entry computed in the current pass (timeStamp) or a failed
// one (-2) from a lower threshold (.cpp:654-656).  Additionally an entry with a real path computed
// earlier in the SAME outer iteration under a lower positive threshold is bit-identical to what a
// recomputation would give as long as no constraint on its path is at or below the new threshold:
// raising minAbundance only removes candidates or lowers their value, abundances do not change inside
// an iteration, and the uncovered-edge rule is the same for every positive threshold -- so the winner
// of the left fold (first maximum) is unchanged.  This saves most of the per-threshold recomputation.
PSC_DEV bool dp_memo_hit(const DpState &st, long long mi, int strand)
{
	const DpMemo *me = st.d.ent + mi;
	if (st.d.eepoch[mi] != st.epoch || me->cover == -1 || me->strand != strand) return false;
	if (me->timeStamp == st.timeStamp) return true;
	if (me->minAbundance < st.minAbundance && me->cover == -2) return true;
	return !st.forAbundance && me->mode == 0 && me->iter == st.iter && me->cover >= 0 && me->minAbundance > 0 &&
	       st.minAbundance > me->minAbundance && me->beta > st.minAbundance;
}

// stable sort of an extension list by .last (.cpp:781-796; glibc qsort == merge sort).  The keys are
// subexon indices just past the key's end (constraints are truncated to a few subexons): a stable
// counting sort over the few distinct values, one ballot per (value, 32 elements); wide key ranges fall
// back to ranking every element against all others.
PSC_DEV void dp_sort_ext(const DpState &st, int *src, int extCnt, bool hasRoom)
{
	const Coop &co = st.co;
	const DpScratch &d = st.d;
	if (extCnt <= 1) return;
	if (co.n > 1 && extCnt <= co.n) {
		// one element per lane: place the distinct keys in ascending order, ties by original position
		bool have = co.lane < extCnt;
		int x = have ? src[co.lane] : -1;
		int key = have ? d.tcl[x] : 0x7fffffff;
		int rank = 0, base = 0;
		for (;;) {
			int kv = co.min(key);
			if (kv == 0x7fffffff) break;
			uint32_t m = co.ballot(key == kv);
			if (key == kv) {
				rank = base + co.rank_in(m);
				key = 0x7fffffff;
			}
			base += popc32(m);
		}
		co.sync();
		if (have) src[rank] = x;
		co.sync();
		return;
	}
	int *tmp = src + extCnt;
	if (hasRoom) {
		int kmin = 0x7fffffff, kmax = -1;
		for (int a = co.lane; a < extCnt; a += co.n) {
			int k = d.tcl[src[a]];
			if (k < kmin) kmin = k;
			if (k > kmax) kmax = k;
		}
		kmin = co.min(kmin);
		kmax = co.max(kmax);
		if (kmax - kmin < 24) {
			int base = 0;
			for (int kv = kmin; kv <= kmax; ++kv)
				for (int b0 = 0; b0 < extCnt; b0 += co.n) {
					int a = b0 + co.lane;
					int x = a < extCnt ? src[a] : -1;
					bool is = a < extCnt && d.tcl[x] == kv;
					uint32_t m = co.ballot(is);
					if (is) tmp[base + co.rank_in(m)] = x;
					base += popc32(m);
				}
		} else {
			for (int a = co.lane; a < extCnt; a += co.n) {
				int x = src[a], xl = d.tcl[x], r = 0;
				for (int b = 0; b < extCnt; ++b) {
					int yl = d.tcl[src[b]];
					if (yl < xl || (yl == xl && b < a)) ++r;
				}
				tmp[r] = x;
			}
		}
		co.sync();
		for (int a = co.lane; a < extCnt; a += co.n) src[a] = tmp[a];
		co.sync();
	} else {
		if (co.lane == 0)
			for (int a = 1; a < extCnt; ++a) {
				int x = src[a], xl = d.tcl[x], b = a - 1;
				while (b >= 0 && d.tcl[src[b]] > xl) { src[b + 1] = src[b]; --b; }
				src[b + 1] = x;
			}
		co.sync();
	}
}

} // namespace pscdev
#include "k_dp_walk.cuh"
namespace pscdev {
// what dp_pick needs from a SolveSubTranscript engine: the result of one root call
struct DpRootRes {
	double cover;
	const uint64_t *bits;
	int first, last;
};
} // namespace pscdev
#include "k_dp2.cuh"
namespace pscdev {


// copy a memo entry into one of the two "best so far" registers kept in work[] (SetDpContent)
struct DpBest {
	double cover;
	int first, last;
	uint64_t *bits;
};
PSC_DEV void dp_take(const DpState &st, DpBest &dst, const uint64_t *bits, double cover, int first, int last)
{
	st.co.sync();
	if (st.co.lane == 0)
		for (int w = 0; w < st.W; ++w) dst.bits[w] = bits[w];
	dst.cover = cover;
	dst.first = first;
	dst.last = last;
	st.co.sync();
}

// the iterative walk with the f1/f2/hash memo (k_dp_walk.cuh)
struct DpEngineWalk {
	static PSC_DEV DpRootRes root(DpState &st, int i)
	{
		long long r = dp_solve_root(st, i);
		DpRootRes o;
		o.cover = st.d.ent[r].cover;
		o.bits = st.d.ebits + r * st.W;
		o.first = st.d.ent[r].first;
		o.last = st.d.ent[r].last;
		return o;
	}
	static PSC_DEV void abundances_changed(DpState &) {}
};

// PickTranscriptsByDP (TranscriptDecider.cpp:860-1201) on the constraints currently loaded in
// st.d.tc*; the new transcripts are appended to st.d.tx* and their count returned.
template <class Engine> PSC_DEVN int dp_pick(DpState &st, int iterBound)
{
	const Coop &co = st.co;
	const CompView &v = st.v;
	DpScratch &d = st.d;
	int W = st.W, n = v.n, tcCnt = d.tcCnt;
	st.timeStamp = 1;
	st.iter = 0;
	++st.epoch;
	// tcEnd[q] = number of constraints starting at or before subexon q (they are sorted by .first)
	for (int q = co.lane; q < n; q += co.n) {
		int lo = 0, hi = tcCnt;
		while (lo < hi) {
			int mid = (lo + hi) >> 1;
			if (d.tcf[mid] <= q) lo = mid + 1;
			else hi = mid;
		}
		d.tcEnd[q] = lo;
	}
	co.sync();
	// ---- uncovered edges (.cpp:890-947): lanes over edges ----
	{
		int e0 = v.noff[0];
		for (int i = 0; i < n; ++i)
			for (int e = v.noff[i] + co.lane; e < v.noff[i + 1]; e += co.n) {
				int nb = v.nidx[e];
				bool covered = false;
				int end = dp_tc_end(st, 0, i);
				for (int l = 0; l < end && !covered; ++l) {
					if (d.tcl[l] < i) continue;
					const uint64_t *b = d.tcb + (long long)l * W;
					if (!(bt_test(b, i) && bt_test(b, nb))) continue;
					bool between = false;
					for (int q = i + 1; q < nb; ++q)
						if (bt_test(b, q)) { between = true; break; }
					if (!between) covered = true;
				}
				d.uncov[e - e0] = covered ? 0 : 1;
			}
		co.sync();
	}
	// ---- abundance of the most abundant path (.cpp:951-966) ----
	double maxAbundance = -1;
	st.forAbundance = true;
	st.minAbundance = 0;
	for (int i = 0; i < n; ++i)
		if (st.cs[i]) {
			long long tq = st.prof ? clock64() : 0;
			DpRootRes r = Engine::root(st, i);
			if (st.prof) st.cycRoot += clock64() - tq;
			double c = r.cover;
			if (c > maxAbundance) maxAbundance = c;
		}
	st.forAbundance = false;
	for (int k = co.lane; k < d.cCap; k += co.n) d.cvalid[k] = 0;
	co.sync();
	DpBest maxCover, best;
	maxCover.bits = d.work;
	best.bits = d.work + W;
	uint64_t *sub = d.work + 2 * W;
	int iterCnt = 0, txCnt = 0;
	for (;;) {
		double bestScore = -1;
		st.minAbundance = 0;
		++st.iter; // constraint abundances change between iterations: memo entries do not carry over
		best.cover = -1;
		for (;;) {
			// cached result for this minAbundance? (.cpp:996-1000)
			int hit = -1;
			for (int k0 = 0; k0 < d.cCap && hit < 0; k0 += co.n) {
				int k = k0 + co.lane;
				uint32_t m = co.ballot(k < d.cCap && d.cvalid[k] && d.ckey[k] == st.minAbundance);
				if (m) hit = k0 + ctz32(m);
			}
			if (hit >= 0) {
				dp_take(st, maxCover, d.cbits + (long long)hit * W, d.cent[hit].cover, d.cent[hit].first, d.cent[hit].last);
			} else {
				maxCover.cover = -1;
				++st.timeStamp;
				++st.nLevels;
				for (int i = 0; i < n; ++i) {
					if (!st.cs[i]) continue;
					long long tq = st.prof ? clock64() : 0;
					DpRootRes r = Engine::root(st, i);
					if (st.prof) st.cycRoot += clock64() - tq;
					double c = r.cover;
					if (c > maxCover.cover && c > 0) dp_take(st, maxCover, r.bits, c, r.first, r.last);
				}
				if (maxCover.cover == -1) break;
				int slot = -1;
				for (int k0 = 0; k0 < d.cCap && slot < 0; k0 += co.n) {
					int k = k0 + co.lane;
					uint32_t m = co.ballot(k < d.cCap && !d.cvalid[k]);
					if (m) slot = k0 + ctz32(m);
				}
				if (slot < 0) {
					*st.err = 22;
					slot = 0;
				}
				co.sync();
				if (co.lane == 0) {
					d.cvalid[slot] = 1;
					d.ckey[slot] = st.minAbundance;
					d.cent[slot].cover = maxCover.cover;
					d.cent[slot].first = maxCover.first;
					d.cent[slot].last = maxCover.last;
					for (int w = 0; w < W; ++w) d.cbits[(long long)slot * W + w] = maxCover.bits[w];
				}
				co.sync();
			}
			// bottleneck abundance of the max-cover path (.cpp:1028-1055)
			double mn = 1e300;
			int anyc = 0;
			for (int i = co.lane; i < tcCnt; i += co.n)
				if (compat(maxCover.bits, maxCover.first, maxCover.last, false, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 1) {
					anyc = 1;
					if (d.tcnorm[i] < mn) mn = d.tcnorm[i];
				}
			mn = co.min(mn);
			anyc = co.max(anyc);
			double min = anyc ? mn : -1;
			if (st.minAbundance == 0) {
				int prev = -1;
				for (int q = maxCover.first; q <= maxCover.last; ++q) {
					if (!bt_test(maxCover.bits, q)) continue;
					if (prev >= 0 && dp_edge_uncovered(st, prev, q)) { min = 1e-6; break; }
					prev = q;
				}
			}
			double score = compute_score(maxCover.cover, 1.0, min, maxAbundance, 0);
			if (bestScore == -1 || score > bestScore) {
				bestScore = score;
				dp_take(st, best, maxCover.bits, maxCover.cover, maxCover.first, maxCover.last);
			} else if (score < bestScore) {
				if (compute_score(maxCover.cover, 1.0, maxAbundance, maxAbundance, 0) < bestScore) break;
			}
			st.minAbundance = min;
		}
		if (best.cover == -1) break;
		// ---- assign the constraints of the best path (.cpp:1074-1152) ----
		long long tqa = clock64();
		int coveredCnt = 0;
		double update = 1e300;
		int anyu = 0;
		for (int b0 = 0; b0 < tcCnt; b0 += co.n) {
			int i = b0 + co.lane;
			bool is = i < tcCnt && compat(best.bits, best.first, best.last, false, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 1;
			uint32_t m = co.ballot(is);
			if (is) {
				d.covered[coveredCnt + co.rank_in(m)] = i;
				if (d.tcab[i] > 0) { anyu = 1; if (d.tcab[i] < update) update = d.tcab[i]; }
			}
			coveredCnt += popc32(m);
		}
		update = co.min(update);
		anyu = co.max(anyu);
		if (!anyu) update = -1;
		update *= (1 + iterCnt / 50); // Q7: integer division
		co.sync();
		if (txCnt < d.txCap) {
			if (co.lane == 0) {
				for (int w = 0; w < W; ++w) d.txb[(long long)txCnt * W + w] = best.bits[w];
				d.txf[txCnt] = best.first;
				d.txl[txCnt] = best.last;
			}
			++txCnt;
		} else
			*st.err = 23;
		// subtract; a constraint that runs dry invalidates the cached paths containing it
		bool ranDry = false;
		for (int q = 0; q < coveredCnt; ++q) {
			int j = d.covered[q];
			double ab = d.tcab[j];
			if (ab > 0) {
				double tmp = ab > update ? update : ab;
				ab -= tmp;
				if (ab <= 0) {
					ranDry = true;
					for (int k = co.lane; k < d.cCap; k += co.n)
						if (d.cvalid[k] && compat(d.cbits + (long long)k * W, d.cent[k].first, d.cent[k].last, false,
						                          d.tcb + (long long)j * W, d.tcf[j], d.tcl[j], W) == 1)
							d.cvalid[k] = 0;
				}
			}
			if (ab < 0) ab = 0;
			co.sync();
			if (co.lane == 0) d.tcab[j] = ab;
			co.sync();
		}
		if (ranDry) Engine::abundances_changed(st);
		st.cycAssign += clock64() - tqa;
		++iterCnt;
		++st.nIters;
		if (iterCnt >= iterBound) break;
	}
	(void)sub;
	return txCnt;
}

// CompSortTranscripts (TranscriptDecider.hpp:360-374): ascending first, then the vector WITH the
// first differing bit first
PSC_DEV bool tx_row_less(const uint64_t *a, int fa, const uint64_t *b, int fb, int W)
{
	if (fa != fb) return fa < fb;
	for (int w = 0; w < W; ++w)
		if (a[w] != b[w]) return brev64(a[w]) > brev64(b[w]);
	return false;
}

// CoalesceSameTranscripts (TranscriptDecider.cpp:303-328) on the candidate table of component c
PSC_DEV int dp_coalesce(DpState &st, uint64_t *cb, int *cf, int *cl, int A)
{
	const Coop &co = st.co;
	DpScratch &d = st.d;
	int W = st.W;
	if (A <= 1) return A;
	if (d.sort64 && next_pow2(A) <= d.sort64Cap && st.v.n < 4096 && A <= 4096) {
		// sort (52-bit prefix | 12-bit row index) words in shared memory: the prefix is `first` and the 40 subexons
		// after it, complemented so that the row WITH the earlier bit comes first (CompSortTranscripts); only rows
		// with equal prefixes are compared in full
		for (int i = co.lane; i < A; i += co.n) {
			const uint64_t *r = cb + (long long)i * W;
			int f = cf[i];
			uint64_t next40 = 0; // bits f+1 .. f+40, the earliest subexon in the top position
			for (int q = 0; q < 40; ++q) {
				int b = f + 1 + q;
				next40 = (next40 << 1) | (uint64_t)((b < st.v.n && bt_test(r, b)) ? 0 : 1);
			}
			d.sort64[i] = ((uint64_t)f << 52) | (next40 << 12) | (uint64_t)i;
		}
		co.sync();
		coop_sort64(co, d.sort64, A, [&](uint64_t x, uint64_t y) {
			if ((x >> 12) != (y >> 12)) return x < y;
			int a = (int)(x & 4095), b = (int)(y & 4095);
			const uint64_t *ra = cb + (long long)a * W, *rb = cb + (long long)b * W;
			for (int w = 0; w < W; ++w)
				if (ra[w] != rb[w]) return brev64(ra[w]) > brev64(rb[w]);
			return a < b;
		});
		for (int i = co.lane; i < A; i += co.n) d.sortidx[i] = (int)(d.sort64[i] & 4095);
		co.sync();
	} else {
	for (int i = co.lane; i < A; i += co.n) d.sortidx[i] = i;
	co.sync();
	coop_sort(co, d.sortidx, A, [&](int a, int b) {
		int fa = cf[a], fb = cf[b];
		if (fa != fb) return fa < fb;
		const uint64_t *ra = cb + (long long)a * W, *rb = cb + (long long)b * W;
		for (int w = 0; w < W; ++w)
			if (ra[w] != rb[w]) return brev64(ra[w]) > brev64(rb[w]); // CompSortTranscripts, see tx_row_less
		return a < b; // equal rows: keep the order of arrival
	});
	}
	// unique rows -> second buffer, then back
	int K = 0;
	for (int b0 = 0; b0 < A; b0 += co.n) {
		int p = b0 + co.lane;
		bool head = false;
		int me = -1;
		if (p < A) {
			me = d.sortidx[p];
			if (p == 0) head = true;
			else {
				int pv = d.sortidx[p - 1];
				head = false;
				for (int w = 0; w < W; ++w)
					if (cb[(long long)me * W + w] != cb[(long long)pv * W + w]) { head = true; break; }
			}
		}
		uint32_t m = co.ballot(head);
		if (head) {
			int k = K + co.rank_in(m);
			for (int w = 0; w < W; ++w) d.c2b[(long long)k * W + w] = cb[(long long)me * W + w];
			d.c2f[k] = cf[me];
			d.c2l[k] = cl[me];
		}
		K += popc32(m);
	}
	co.sync();
	for (int k = co.lane; k < K; k += co.n) {
		for (int w = 0; w < W; ++w) cb[(long long)k * W + w] = d.c2b[(long long)k * W + w];
		cf[k] = d.c2f[k];
		cl[k] = d.c2l[k];
	}
	co.sync();
	return K;
}

// AugmentTranscripts(limit, extend = false) (TranscriptDecider.cpp:1205-1365)
PSC_DEV int dp_augment(DpState &st, uint64_t *cb, int *cf, int *cl, int A, int limit, long long cap)
{
	const Coop &co = st.co;
	const CompView &v = st.v;
	int W = st.W;
	if (A >= limit) return A;
	int size = A;
	bool full = false;
	co.sync();
	// The reference appends, row by row, the suffixes that start at a possible start and then the prefixes that end
	// at a possible end, and stops the moment `limit` rows exist.  Same rows in the same order here, but counted per
	// source row first (lanes over rows), placed by a running offset, and cut at the limit.
	const int room = limit - A; // > 0
	int base = 0;               // rows appended by the source rows before the current batch
	for (int b0 = 0; b0 < size && base < room; b0 += co.n) {
		const int i = b0 + co.lane;
		int nSuf = 0, nPre = 0;
		if (i < size) {
			const uint64_t *src = cb + (long long)i * W;
			const int tf = cf[i], tl = cl[i];
			for (int q = tf + 1; q <= tl; ++q)
				if (bt_test(src, q) && st.cs[q]) ++nSuf; // bit tf is the first set bit: suffixes start at later ones
			for (int q = tf; q < tl; ++q)
				if (bt_test(src, q) && st.ce[q]) ++nPre;
		}
		const int mine = nSuf + nPre;
		// exclusive prefix over the lanes of this batch
		int off = 0, tot = 0;
		for (int l = 0; l < co.n; ++l) {
			int v = co.bcast(mine, l);
			if (l < co.lane) off += v;
			tot += v;
		}
		if (i < size && mine > 0) {
			const uint64_t *src = cb + (long long)i * W;
			const int tf = cf[i], tl = cl[i];
			int k = base + off; // index among the appended rows
			for (int q = tf + 1; q <= tl && k < room; ++q)
				if (bt_test(src, q) && st.cs[q]) {
					if ((long long)A + k >= cap) { *st.err = 24; break; }
					uint64_t *dst = cb + (long long)(A + k) * W;
					for (int w = 0; w < W; ++w) dst[w] = src[w] & keep_mask(w, q, tl);
					cf[A + k] = q;
					cl[A + k] = tl;
					++k;
				}
			for (int q = tf; q < tl && k < room; ++q)
				if (bt_test(src, q) && st.ce[q]) {
					if ((long long)A + k >= cap) { *st.err = 24; break; }
					uint64_t *dst = cb + (long long)(A + k) * W;
					for (int w = 0; w < W; ++w) dst[w] = src[w] & keep_mask(w, tf, q);
					cf[A + k] = tf;
					cl[A + k] = q;
					++k;
				}
		}
		base += tot;
	}
	if (base >= room) {
		base = room;
		full = true;
	}
	A += base;
	if ((long long)A > cap) A = (int)cap;
	co.sync();
	if (full) return A; // the reference returns before the final coalesce when the limit is hit
	return dp_coalesce(st, cb, cf, cl, A);
}

// -----------------------------------------------------------------------------------------------
// The DP branch of Solve (TranscriptDecider.cpp:3139-3265), split in two so that the per-sample
// PickTranscriptsByDP calls of one component run concurrently:
//   dp_unit_pick : one cooperative group per (DP component, sample).  The call only depends on the
//                  sample's constraints and on iterBound; a smaller iterBound yields a PREFIX of the
//                  same transcript list (the loop of .cpp:983-1165 just stops earlier), so every unit
//                  runs with the largest bound it can get and the merge takes the prefix it needs.
//   dp_merge     : one group per component, samples in the reference's order (descending number of
//                  constraints, stable): append, CoalesceSameTranscripts, AugmentTranscripts.
// -----------------------------------------------------------------------------------------------
struct DpUnit {
	int comp, sample;
	int iterBound; // C of the sample for the first (largest) sample of the component, else 100
	int stride;    // 0: constraints as they are (or truncated); > 0: DownsampleConstraintsFrom stride
	long long scratch_off; // DpScratch of the unit (memo part zeroed by the host)
	long long tx_off;      // bytes into the transcript-list arena
};

struct DpTxList {
	uint64_t *b;
	int *f, *l;
};
__host__ __device__ inline size_t dp_txlist_layout(DpTxList *o, char *base, int cap, int W)
{
	Carver cv(base);
	DpTxList t;
	t.b = cv.take<uint64_t>((size_t)cap * W + 1);
	t.f = cv.take<int>((size_t)cap + 1);
	t.l = cv.take<int>((size_t)cap + 1);
	if (o) *o = t;
	return (cv.off + 15) & ~(size_t)15;
}

// The control state of the explicit-stack machine (path, frames, per-frame best path, the shared
// transcript buffer) is what every step reads and writes; when it fits the warp's slice of shared
// memory it lives there instead of in the global scratch.
PSC_DEV void dp_use_shared(DpScratch &d, int n, int W, char *sm, size_t smBytes)
{
	if (!sm) return;
	Carver cv(sm);
	int *P = cv.take<int>((size_t)n + 4);
	int *fr = cv.take<int>((size_t)PSC_FR_INTS_TOTAL(n));
	double *fdcover = cv.take<double>((size_t)n + 3);
	double *fdbeta = cv.take<double>((size_t)n + 3);
	int *fdmeta = cv.take<int>((size_t)(n + 3) * 4);
	uint64_t *fdbits = cv.take<uint64_t>((size_t)(n + 3) * W);
	uint64_t *fdkey = cv.take<uint64_t>((size_t)(n + 3) * W);
	uint64_t *pbits = cv.take<uint64_t>((size_t)(n + 4) * W);
	uint64_t *work = cv.take<uint64_t>((size_t)8 * W);
	if (cv.off > smBytes) return;
	d.P = P;
	d.fr = fr;
	d.fdcover = fdcover;
	d.fdbeta = fdbeta;
	d.fdmeta = fdmeta;
	d.fdbits = fdbits;
	d.fdkey = fdkey;
	d.pbits = pbits;
	d.work = work;
}

// the constraints PickTranscriptsByDP sees: truncated to their first maxDp subexons
// (TruncateConstraintsCoverFrom, Constraints.hpp:400-434) or every stride-th one
// (DownsampleConstraintsFrom, Constraints.hpp:362-398)
PSC_DEV void dp_load_constraints(DpState &st, const TableView &t, int unitStride, int maxDp)
{
	const Coop &co = st.co;
	const int W = st.W;
	if (maxDp > 0) {
		// TruncateConstraintsCoverFrom (Constraints.hpp:400-434)
		for (int k = co.lane; k < t.C; k += co.n) {
			const uint64_t *src = t.bits + (long long)k * W;
			uint64_t *dst = st.d.tcb + (long long)k * W;
			for (int w = 0; w < W; ++w) dst[w] = 0;
			int taken = 0, last = t.first[k];
			for (int q = t.first[k]; q <= t.last[k] && taken < maxDp; ++q)
				if (bt_test(src, q)) { bt_set(dst, q); last = q; ++taken; }
			st.d.tcf[k] = t.first[k];
			st.d.tcl[k] = last;
			st.d.tcsup[k] = t.support[k];
			st.d.tcnorm[k] = t.norm[k];
			st.d.tcab[k] = t.norm[k];
		}
		st.d.tcCnt = t.C;
	} else {
		int stride = unitStride > 0 ? unitStride : 1; // DownsampleConstraintsFrom (Constraints.hpp:362-398)
		int cnt = (t.C + stride - 1) / stride;
		for (int k = co.lane; k < cnt; k += co.n) {
			int src = k * stride;
			for (int w = 0; w < W; ++w) st.d.tcb[(long long)k * W + w] = t.bits[(long long)src * W + w];
			st.d.tcf[k] = t.first[src];
			st.d.tcl[k] = t.last[src];
			st.d.tcsup[k] = t.support[src];
			st.d.tcnorm[k] = t.norm[src];
			st.d.tcab[k] = t.norm[src];
		}
		st.d.tcCnt = cnt;
	}
}

PSC_DEVN void dp_unit_pick(const Coop &co, const GraphDev &g, const SolveDev &sv, const DpUnit &u, char *scratchBase,
                           char *txBase, int *n_tx_out, char *sm, size_t smBytes, long long *dbg)
{
	DpState st;
	st.nStates = st.nLevels = st.nIters = st.nFrames = st.nRoots = st.cycRoot = st.cycAssign = 0;
	long long t0 = clock64();
	st.co = co;
	st.v = comp_view(g, u.comp);
	const CompView &v = st.v;
	int c = u.comp;
	int E = v.noff[v.n] - v.noff[0];
	TableView t = table_view(sv.samples[u.sample], c);
	int Ccap = t.C > 100 ? t.C : 100;
	dp_scratch_layout(&st.d, scratchBase + u.scratch_off, v.n, E, Ccap, 1);
	dp_use_shared(st.d, v.n, v.W, sm, smBytes);
	{
		long long nEnt = dp_memo_entries(v.n);
		for (long long k = co.lane; k < nEnt; k += co.n) st.d.eepoch[k] = 0;
	}
	DpTxList out;
	dp_txlist_layout(&out, txBase + u.tx_off, u.iterBound, v.W);
	// the unit's transcript list is written straight into its persistent slot
	st.d.txb = out.b;
	st.d.txf = out.f;
	st.d.txl = out.l;
	st.d.txCap = u.iterBound;
	st.cs = sv.can_start + v.base;
	st.ce = sv.can_end + v.base;
	st.W = v.W;
	st.epoch = 0;
	st.prof = dbg != nullptr;
	st.err = sv.error_flag;
	st.z = nullptr;
	st.forAbundance = false;
	st.minAbundance = 0;
	st.timeStamp = 0;
	dp_load_constraints(st, t, u.stride, sv.prm.maxDpConstraintSize);
	co.sync();
	int nt = dp_pick<DpEngineWalk>(st, u.iterBound);
	if (co.lane == 0) *n_tx_out = nt;
	if (dbg && co.lane == 0) {
		dbg[0] = clock64() - t0;
		dbg[1] = st.nStates + ((st.cycRoot >> 20) << 32);
		dbg[2] = st.nLevels + ((st.cycAssign >> 20) << 32);
		dbg[3] = st.nIters + (st.nFrames << 12) + (st.nRoots << 40);
	}
	co.sync();
}

// The same unit on the static-program engine (k_dp2.cuh).  `arena` is the calling group's private global working set
// of arenaBytes, `hot` its slice of shared memory (or nullptr).  The unit first runs with its hot tables in shared
// memory; if they outgrow the slice it starts over with everything in the arena; a unit whose key / frame / list
// tables outgrow that too reports *n_tx_out = -1 and is redone by the walk engine.
PSC_DEVN void dp2_unit_pick(const Coop &co, const GraphDev &g, const SolveDev &sv, const DpUnit &u, char *arena, size_t arenaBytes,
                            char *hot, size_t hotBytes, char *txBase, int *n_tx_out, long long *dbg, int roomy = 0)
{
	DpState st;
	st.nStates = st.nLevels = st.nIters = st.nFrames = st.nRoots = st.cycRoot = st.cycAssign = 0;
	long long t0 = clock64();
	st.co = co;
	st.v = comp_view(g, u.comp);
	const CompView &v = st.v;
	int E = v.noff[v.n] - v.noff[0];
	TableView t = table_view(sv.samples[u.sample], u.comp);
	int Ccap = t.C > 100 ? t.C : 100;
	Dp2 z;
	int nt = -1;
	int inShared = 0;
	// shared memory is tried only when the slice holds the tables of a typical unit of this size (~1.4 n keys on the
	// bench workload): an attempt that overflows is thrown away
	const bool trySmem = !roomy && hot && dp2_hot_fixed(v.n, v.W) + (size_t)(3 * v.n / 2 + 8) * dp2_hot_per_key(v.W, DP2_FR_PER_KEY_SMEM) <= hotBytes;
	for (int attempt = trySmem ? 0 : 1; attempt < 2; ++attempt) {
		char *h = attempt == 0 ? hot : nullptr;
		size_t used = dp2_layout(&st.d, &z, arena, arenaBytes, h, hotBytes, v.n, E, Ccap, 0, roomy ? DP2_FR_PER_KEY_ROOMY : DP2_FR_PER_KEY_GLOBAL);
		if (used > arenaBytes || z.maxKeys < 8 || (h && z.hotBytesUsed > hotBytes)) continue; // does not even hold the fixed part
		inShared = h != nullptr;
		for (int k = co.lane; k < v.n + 1; k += co.n) z.key1[k] = -1;
		for (int k = co.lane; k < z.hashCap; k += co.n) z.khash[k] = -1;
		DpTxList out;
		dp_txlist_layout(&out, txBase + u.tx_off, u.iterBound, v.W);
		st.d.txb = out.b;
		st.d.txf = out.f;
		st.d.txl = out.l;
		st.d.txCap = u.iterBound;
		st.cs = sv.can_start + v.base;
		st.ce = sv.can_end + v.base;
		st.W = v.W;
		st.epoch = 0;
		st.prof = dbg != nullptr;
		st.err = sv.error_flag;
		st.z = &z;
		st.forAbundance = false;
		st.minAbundance = 0;
		st.timeStamp = 0;
		dp_load_constraints(st, t, u.stride, sv.prm.maxDpConstraintSize);
		co.sync();
		nt = dp_pick<DpEngineProgram>(st, u.iterBound);
		co.sync();
		if (!z.overflow) break;
		nt = -1;
	}
	if (co.lane == 0) *n_tx_out = nt;
	if (dbg && co.lane == 0) {
		dbg[0] = clock64() - t0;
		dbg[1] = st.nStates + ((st.cycRoot >> 20) << 32);
		dbg[2] = st.nLevels + ((st.cycAssign >> 20) << 32);
		dbg[3] = st.nIters + (st.nFrames << 12) + (st.nRoots << 40);
		dbg[4] = z.nKeys;
		dbg[5] = z.nFrames;
		dbg[6] = z.nList;
		dbg[7] = nt < 0 ? -1 : (inShared ? -2 - (long long)z.maxKeys : z.maxKeys);
	}
	co.sync();
}

struct DpMergeScratch {
	int *sortidx;
	uint64_t *c2b;
	int *c2f, *c2l;
};
__host__ __device__ inline size_t dp_merge_layout(DpMergeScratch *o, char *base, int W, long long candCap)
{
	Carver cv(base);
	DpMergeScratch s;
	long long p2 = 1;
	while (p2 < candCap + 1) p2 <<= 1;
	s.sortidx = cv.take<int>((size_t)p2 + 1);
	s.c2b = cv.take<uint64_t>((size_t)candCap * W + 1);
	s.c2f = cv.take<int>((size_t)candCap + 1);
	s.c2l = cv.take<int>((size_t)candCap + 1);
	if (o) *o = s;
	return (cv.off + 15) & ~(size_t)15;
}

// units of component c are u0 .. u0+S-1, already in the reference's visiting order
PSC_DEVN void dp_merge(const Coop &co, const GraphDev &g, const SolveDev &sv, int c, const DpUnit *units, int u0,
                       const int *n_tx, const char *txBase, char *scratch, uint64_t *sort64, int sort64Cap)
{
	DpState st;
	st.co = co;
	st.v = comp_view(g, c);
	const CompView &v = st.v;
	int S = sv.prm.S, W = v.W;
	long long row0 = sv.cand_row_off[c];
	long long cap = sv.cand_row_off[c + 1] - row0;
	DpMergeScratch ms;
	dp_merge_layout(&ms, scratch, W, cap);
	st.d.sortidx = ms.sortidx;
	st.d.c2b = ms.c2b;
	st.d.c2f = ms.c2f;
	st.d.c2l = ms.c2l;
	st.d.sort64 = sort64;
	st.d.sort64Cap = sort64Cap;
	st.cs = sv.can_start + v.base;
	st.ce = sv.can_end + v.base;
	st.W = W;
	st.err = sv.error_flag;
	uint64_t *cb = sv.cand_bits + sv.cand_bits_off[c];
	int *cf = sv.cand_first + row0, *cl = sv.cand_last + row0;
	int A = 0;
	for (int k = 0; k < S; ++k) {
		const DpUnit &u = units[u0 + k];
		int oi = S - 1 - k; // position in the ascending order the reference counts with
		int bound = u.iterBound;
		if (oi < S - 10 && A > 1000) bound = 10;
		if (u.stride > 0 && bound <= 10) continue; // down-sampled samples with a tiny budget are skipped (.cpp:3228)
		int nt = n_tx[u0 + k];
		if (nt > bound) nt = bound;
		DpTxList src;
		dp_txlist_layout(&src, const_cast<char *>(txBase) + u.tx_off, u.iterBound, W);
		if (A + nt > cap) {
			*st.err = 25;
			nt = (int)(cap - A);
		}
		for (int q = co.lane; q < nt; q += co.n) {
			for (int w = 0; w < W; ++w) cb[(long long)(A + q) * W + w] = src.b[(long long)q * W + w];
			cf[A + q] = src.f[q];
			cl[A + q] = src.l[q];
		}
		A += nt;
		co.sync();
		A = dp_coalesce(st, cb, cf, cl, A);
		A = dp_augment(st, cb, cf, cl, A, 1000, cap);
	}
	for (int k = co.lane; k < A; k += co.n) {
		sv.cand_corr[row0 + k] = 0;
		sv.cand_sup1[row0 + k] = 0;
		sv.cand_sup2[row0 + k] = 0;
	}
	if (co.lane == 0) sv.n_cand[c] = A;
	co.sync();
}

} // namespace pscdev
