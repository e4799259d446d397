// k_solve.cuh -- TranscriptDecider::Solve on the device (TranscriptDecider.cpp:2874-3460).
//
// Work units: one cooperative group (Coop, 1 or 32 lanes) per component for the cross-sample
// steps and per (component, sample) for the per-sample steps; kernel boundaries stand where the
// reference has its three cross-sample reductions.  All working storage is
// carved out of per-unit scratch arenas in HBM that the host sizes from the table sizes.
#pragma once
#include "psc_internal.h"
#include "pow_quarter.h"

namespace pscdev {

#define PSC_USE_DP 200000

struct SolveParams {
	double fpkmFraction, classifierThreshold, minDepth;
	int maxDpConstraintSize, S;
	int readLen, fragLen, fragStdev, matePaired;
};

// ---- bump allocator used identically by the host (sizing, base == 0) and the device ----
struct Carver {
	char *base;
	size_t off;
	__host__ __device__ Carver(char *b) : base(b), off(0) {}
	template <class T> __host__ __device__ T *take(size_t n)
	{
		off = (off + 15) & ~(size_t)15;
		T *p = (T *)(base + off);
		off += n * sizeof(T);
		return p;
	}
};

// ---- per-component state that persists through the whole solve ----
struct SolveDev {
	SolveParams prm;
	const SampleDev *samples;
	// prologue results
	uint8_t *can_start, *can_end; // [n_se]
	int *edge_support;            // [n_next]  #samples supporting the intron edge (subexonChainSupport)
	int *cs00;                    // [n_comp]  the {0,0} entry an empty chain list leaves behind
	int *at_cnt, *use_dp;         // [n_comp]
	int *def_gene;                // [2*n_comp]
	// candidate transcripts (alltranscripts)
	const long long *cand_row_off;  // [n_comp+1]
	const long long *cand_bits_off; // [n_comp+1] in words
	int *n_cand;                    // [n_comp]
	uint64_t *cand_bits;
	int *cand_first, *cand_last;
	int8_t *cand_corr;          // correlationScore of the candidate (1 for single-subexon enumerated paths)
	int *cand_sup1, *cand_sup2; // #samples that picked it in round 1 / kept it after refinement
	int *error_flag;
};

// per-unit results that live from the pick kernel to the emit kernel
struct UnitPred {
	int *id;
	double *abundance, *fpkm;
	int8_t *corr;
	int *raw_id;
	double *raw_abundance;
};
__host__ __device__ inline size_t unit_pred_layout(UnitPred *o, char *base, int cap)
{
	Carver cv(base);
	int *id = cv.take<int>((size_t)2 * cap);
	double *ab = cv.take<double>((size_t)2 * cap);
	double *fp = cv.take<double>((size_t)2 * cap);
	int8_t *co = cv.take<int8_t>((size_t)2 * cap);
	int *rid = cv.take<int>((size_t)cap);
	double *rab = cv.take<double>((size_t)cap);
	if (o) {
		o->id = id; o->abundance = ab; o->fpkm = fp; o->corr = co; o->raw_id = rid; o->raw_abundance = rab;
	}
	return (cv.off + 15) & ~(size_t)15;
}

// per-unit scratch of PickTranscripts / AbundanceEstimation / RefineTranscripts
#ifndef PSC_OL
#define PSC_OL 12
#endif
struct UnitScratch {
	uint64_t *btable;  // [A * MW]
	double *tAbund, *avgAbund, *covPortion, *coverCnt, *abund; // [A]
	double *pairAbund; // [M]
	int *slot;         // [A] row of the candidate in the csupport pool, -1 if none yet
	double *pool;      // [cap * M]
	int *list;         // [A]
	int *depl;         // [M] pairs depleted in the current iteration
	double *score;     // [A] cached ComputeScore of the candidate, -1: not selectable
	double *contrib;   // [M] this iteration's support of maxtag per pair, -1: pair was already depleted
	int *picked;       // [cap] candidates owning a csupport row, ascending
	int *ocnt;         // [M] number of candidates that ever held support for the pair, -1: more than PSC_OL
	int *olist;        // [M * PSC_OL] those candidates, ascending
	int *reach;        // [32 * n] per-lane scratch of the chain-coverage pass
	double *persist;   // [4] inf, any-compatible flag, maxAbundance: round 1 -> round 2
	uint64_t *tmpbits; // [4 * W]
	uint8_t *flagsA;   // [max(A, n, C)]  small boolean scratch
	// EM / refine (T <= 2*cap transcripts)
	uint64_t *ebtable; // [2*cap * MW]
	double *rho, *sumj, *dtmp; // [2*cap], [M], [2*cap]
	int *tlen, *tgid, *icnt; // [2*cap] each
	double *gMaxFPKM, *gNonOvl, *gMaxCov; // [ng]
	int *gTag, *gRecover; // [ng]
	uint8_t *missing; // [E + 1]
	int *recList;     // [2*cap]
	uint8_t *used;    // [2*cap]
};
__host__ __device__ inline size_t unit_scratch_layout(UnitScratch *o, char *base, int A, int M, int cap, int n, int C,
                                                      int ng, int E)
{
	Carver cv(base);
	int MW = (M + 63) >> 6, W = (n + 63) >> 6;
	UnitScratch s;
	s.btable = cv.take<uint64_t>((size_t)A * MW);
	s.tAbund = cv.take<double>((size_t)A);
	s.avgAbund = cv.take<double>((size_t)A);
	s.covPortion = cv.take<double>((size_t)A);
	s.coverCnt = cv.take<double>((size_t)A);
	s.abund = cv.take<double>((size_t)A);
	s.pairAbund = cv.take<double>((size_t)M);
	s.slot = cv.take<int>((size_t)A);
	s.pool = cv.take<double>((size_t)cap * M);
	s.list = cv.take<int>((size_t)A);
	s.depl = cv.take<int>((size_t)M + 1);
	s.score = cv.take<double>((size_t)A);
	s.contrib = cv.take<double>((size_t)M + 1);
	s.picked = cv.take<int>((size_t)cap + 1);
	s.ocnt = cv.take<int>((size_t)M + 1);
	s.olist = cv.take<int>((size_t)M * PSC_OL + 1);
	s.reach = cv.take<int>((size_t)32 * n);
	s.persist = cv.take<double>(4);
	s.tmpbits = cv.take<uint64_t>((size_t)4 * W);
	int fa = A > n ? A : n;
	if (C > fa) fa = C;
	s.flagsA = cv.take<uint8_t>((size_t)fa + 1);
	s.ebtable = cv.take<uint64_t>((size_t)2 * cap * MW);
	s.rho = cv.take<double>((size_t)2 * cap);
	s.sumj = cv.take<double>((size_t)M);
	s.dtmp = cv.take<double>((size_t)2 * cap);
	s.tlen = cv.take<int>((size_t)2 * cap);
	s.tgid = cv.take<int>((size_t)2 * cap);
	s.icnt = cv.take<int>((size_t)2 * cap);
	s.gMaxFPKM = cv.take<double>((size_t)ng + 1);
	s.gNonOvl = cv.take<double>((size_t)ng + 1);
	s.gMaxCov = cv.take<double>((size_t)ng + 1);
	s.gTag = cv.take<int>((size_t)ng + 1);
	s.gRecover = cv.take<int>((size_t)ng + 1);
	s.missing = cv.take<uint8_t>((size_t)E + 2);
	s.recList = cv.take<int>((size_t)2 * cap);
	s.used = cv.take<uint8_t>((size_t)2 * cap + 1);
	if (o) *o = s;
	return (cv.off + 15) & ~(size_t)15;
}

// a chunk of components being processed together
struct ChunkDev {
	int c0, c1; // components [c0, c1)
	const long long *scratch_off; // [(c1-c0)*S + 1] byte offsets into scratch
	char *scratch;
	const long long *pred_off; // [(c1-c0)*S + 1] byte offsets into pred
	char *pred;
	int *n_pred;      // [(c1-c0)*S]  current size of predTranscripts
	int *n_raw;       // [(c1-c0)*S]
	int *n_exon;      // [(c1-c0)*S]  total exons of the final transcripts
};

// view of one component
struct CompView {
	int c, base, n, W;
	const int *st, *en;
	const int8_t *lt, *rt, *ls, *rs;
	const double *lcls, *rcls;
	const int *gene;
	const int *noff, *nidx, *poff, *pidx; // noff/poff indexed with local subexon index (already shifted by base)
	int baseGene, usedGene;
};
PSC_DEV CompView comp_view(const GraphDev &g, int c)
{
	CompView v;
	v.c = c;
	v.base = g.comp_off[c];
	v.n = g.comp_off[c + 1] - v.base;
	v.W = (v.n + 63) >> 6;
	v.st = g.se_start + v.base;
	v.en = g.se_end + v.base;
	v.lt = g.lt + v.base;
	v.rt = g.rt + v.base;
	v.ls = g.ls + v.base;
	v.rs = g.rs + v.base;
	v.lcls = g.lcls + v.base;
	v.rcls = g.rcls + v.base;
	v.gene = g.gene + v.base;
	v.noff = g.next_off + v.base;
	v.nidx = g.next_idx;
	v.poff = g.prev_off + v.base;
	v.pidx = g.prev_idx;
	v.baseGene = g.comp_base[c];
	v.usedGene = g.comp_used[c];
	return v;
}
PSC_DEV bool same_strand(int a, int b) { return a == 0 || b == 0 || a == b; }

// view of one (sample, component) constraint table
struct TableView {
	int C, M, lo;
	const uint64_t *bits;
	const int *first, *last, *support, *uniq;
	const double *norm;
	const int *pi, *pj, *psup, *puniq, *pec;
	const double *pnorm;
};
PSC_DEV TableView table_view(const SampleDev &s, int c)
{
	TableView t;
	t.lo = s.grp_lo[c];
	t.C = s.n_con[c];
	t.M = s.n_pair[c];
	t.bits = s.con_bits + s.bits_off[c];
	t.first = s.con_first + t.lo;
	t.last = s.con_last + t.lo;
	t.support = s.con_support + t.lo;
	t.uniq = s.con_uniq + t.lo;
	t.norm = s.con_norm + t.lo;
	t.pi = s.pair_i + t.lo;
	t.pj = s.pair_j + t.lo;
	t.psup = s.pair_support + t.lo;
	t.puniq = s.pair_uniq + t.lo;
	t.pec = s.pair_ec + t.lo;
	t.pnorm = s.pair_norm + t.lo;
	return t;
}

// =============================================================================================
// P1: prologue of Solve, one group per component (TranscriptDecider.cpp:2884-3111)
// =============================================================================================
PSC_DEV void solve_prologue(const Coop &co, const GraphDev &g, const SolveDev &sv, int c, int *f /* [n] scratch */)
{
	CompView v = comp_view(g, c);
	int n = v.n;
	uint8_t *cs = sv.can_start + v.base, *ce = sv.can_end + v.base;
	double thr = sv.prm.classifierThreshold;
	int S = sv.prm.S;
	// chain support: every sample marks the intron edges of its well-supported constraints
	// (.cpp:3005-3064).  Edge marks of one sample are idempotent, so lanes may race on them.
	int e0 = v.noff[0], e1 = v.noff[n];
	for (int e = e0 + co.lane; e < e1; e += co.n) sv.edge_support[e] = 0;
	co.sync();
	int cs00 = 0;
	for (int s = 0; s < S; ++s) {
		TableView t = table_view(sv.samples[s], c);
		// pass 1: mark (stamp = s + 1 in the high bits would need atomics; use two passes instead)
		int any = 0;
		for (int k = co.lane; k < t.C; k += co.n) {
			if (t.uniq[k] < 0.95 * t.support[k] || t.support[k] < 3) continue;
			const uint64_t *b = t.bits + (long long)k * v.W;
			int prev = -1;
			for (int q = t.first[k]; q <= t.last[k]; ++q) {
				if (!bt_test(b, q)) continue;
				if (prev >= 0 && v.en[prev] + 1 < v.st[q]) {
					any = 1;
					for (int e = v.noff[prev]; e < v.noff[prev + 1]; ++e)
						if (v.nidx[e] == q) {
							// bit 30 marks "seen in the current sample"
							int old = sv.edge_support[e];
							if (!(old & (1 << 30))) sv.edge_support[e] = old | (1 << 30);
							break;
						}
				}
				prev = q;
			}
		}
		any = co.max(any);
		co.sync();
		for (int e = e0 + co.lane; e < e1; e += co.n) {
			int old = sv.edge_support[e];
			if (old & (1 << 30)) sv.edge_support[e] = (old & ~(1 << 30)) + 1;
		}
		co.sync();
		if (!any) ++cs00; // empty chain: vector::resize(k+1) leaves one {0,0} pair behind (.cpp:3039-3063)
	}
	if (co.lane == 0) {
		sv.cs00[c] = cs00;
		// start / end flags (.cpp:2884-2973), sequential by nature
		for (int i = 0; i < n; ++i) {
			bool st = false, en = false;
			if (v.poff[i + 1] == v.poff[i]) st = true;
			else if (v.lcls[i] < thr && v.lcls[i] != -1 && v.ls[i] != 0) st = true; // Q12: the inner test never fires
			if (v.noff[i + 1] == v.noff[i]) en = true;
			else if (v.rcls[i] < thr && v.rcls[i] != -1 && v.rs[i] != 0) en = true;
			cs[i] = st;
			ce[i] = en;
			if (en)
				for (int j = i - 1; j >= 0; --j) {
					if (v.en[j] + 1 != v.st[j + 1]) break;
					if (ce[j]) { ce[j] = 0; break; }
				}
		}
		for (int i = 0; i < n;) {
			int j, k;
			for (j = i + 1; j < n; ++j)
				if (v.st[j] > v.en[j - 1] + 1) break;
			if (!cs[i]) {
				int cnt[2] = {0, 0};
				for (k = i; k < j; ++k) {
					if (!same_strand(v.rs[k], v.ls[i])) break;
					if (v.ls[k] != 0) ++cnt[(v.ls[k] + 1) / 2];
				}
				if (k < j && cnt[(v.rs[k] + 1) / 2] == 0) cs[i] = 1;
			}
			if (!ce[j - 1]) {
				int cnt[2] = {0, 0};
				for (k = j - 1; k >= i; --k) {
					if (!same_strand(v.ls[k], v.rs[j - 1])) break;
					if (v.rs[k] != 0) ++cnt[(v.rs[k] + 1) / 2];
				}
				if (k >= i && cnt[(v.ls[k] + 1) / 2] == 0) ce[j - 1] = 1;
			}
			i = j;
		}
		// default gene ids (.cpp:2979-3003)
		int d0 = -1, d1 = -1;
		for (int i = 0; i < n; ++i) {
			int gid = v.gene[i];
			if (gid < 0) continue;
			if ((v.rs[i] == -1 || v.ls[i] == -1) && (d0 == -1 || gid < d0)) d0 = gid;
			if ((v.rs[i] == 1 || v.ls[i] == 1) && (d1 == -1 || gid < d1)) d1 = gid;
		}
		if (d0 == -1) d0 = v.baseGene;
		if (d1 == -1) d1 = v.usedGene - 1;
		sv.def_gene[2 * c] = d0;
		sv.def_gene[2 * c + 1] = d1;
		// path counts (SubTranscriptCount, .cpp:284-301): int32 wrap-around (Q4); only subexons
		// reachable from a start are ever visited by the reference's memoised recursion
		for (int i = 0; i < n; ++i) f[i] = 0; // reach flag first
		for (int i = 0; i < n; ++i) {
			if (cs[i]) f[i] = 1;
			if (f[i])
				for (int e = v.noff[i]; e < v.noff[i + 1]; ++e) f[v.nidx[e]] = 1;
		}
		for (int i = n - 1; i >= 0; --i) {
			if (!f[i]) { f[i] = -1; continue; }
			uint32_t r = ce[i] ? 1u : 0u;
			for (int e = v.noff[i]; e < v.noff[i + 1]; ++e) r += (uint32_t)f[v.nidx[e]];
			if (r == 0) r = 1;
			f[i] = (int)r;
		}
		uint32_t cnt = 0;
		for (int i = 0; i < n; ++i)
			if (cs[i]) cnt += (uint32_t)f[i];
		int icnt = (int)cnt;
		bool useDP = false;
		if (icnt <= PSC_USE_DP) {
			for (int i = 0; i < n; ++i)
				if (f[i] > PSC_USE_DP) { useDP = true; break; }
		} else
			useDP = true;
		if (!useDP)
			for (int s = 0; s < S; ++s) {
				double msize = (double)sv.samples[s].n_pair[c], csize = (double)sv.samples[s].n_con[c];
				if (icnt > (csize / msize) * (csize / msize) * n && icnt > PSC_USE_DP / (msize * msize) && icnt > 50) {
					useDP = true;
					break;
				}
			}
		sv.at_cnt[c] = icnt;
		sv.use_dp[c] = useDP ? 1 : 0;
	}
	co.sync();
}

// =============================================================================================
// P2a: EnumerateTranscript (TranscriptDecider.cpp:330-372): depth-first, in index order
// =============================================================================================
PSC_DEV void enumerate_candidates(const Coop &co, const GraphDev &g, const SolveDev &sv, int c, int *stackNode,
                                  int *stackEdge, int *stackStrand)
{
	CompView v = comp_view(g, c);
	if (co.lane != 0) return;
	long long row0 = sv.cand_row_off[c];
	long long cap = sv.cand_row_off[c + 1] - row0;
	uint64_t *bits = sv.cand_bits + sv.cand_bits_off[c];
	const uint8_t *cs = sv.can_start + v.base, *ce = sv.can_end + v.base;
	int W = v.W;
	long long A = 0;
	for (int s0 = 0; s0 < v.n; ++s0) {
		if (!cs[s0]) continue;
		int depth = 0;
		stackNode[0] = s0;
		stackEdge[0] = -1; // -1: node not yet "entered"
		stackStrand[0] = 0;
		while (depth >= 0) {
			int tag = stackNode[depth];
			if (stackEdge[depth] < 0) {
				if (ce[tag]) {
					if (A < cap) {
						uint64_t *b = bits + A * W;
						for (int w = 0; w < W; ++w) b[w] = 0;
						for (int i = 0; i <= depth; ++i) bt_set(b, stackNode[i]);
						sv.cand_first[row0 + A] = stackNode[0];
						sv.cand_last[row0 + A] = tag;
						sv.cand_corr[row0 + A] = depth == 0 ? 1 : 0; // min pairwise correlation: 0 without --ls
						sv.cand_sup1[row0 + A] = 0;
						sv.cand_sup2[row0 + A] = 0;
					} else
						*sv.error_flag = 11;
					++A;
				}
				stackEdge[depth] = v.noff[tag];
			}
			int e = stackEdge[depth];
			if (e >= v.noff[tag + 1]) {
				--depth;
				continue;
			}
			stackEdge[depth] = e + 1;
			int a = v.nidx[e];
			int strand = stackStrand[depth];
			bool intron = v.st[a] > v.en[tag] + 1;
			if (!same_strand(v.rs[tag], strand) && intron) continue;
			if (intron && strand == 0) strand = v.rs[tag];
			++depth;
			stackNode[depth] = a;
			stackEdge[depth] = -1;
			stackStrand[depth] = strand;
		}
	}
	sv.n_cand[c] = (int)(A < cap ? A : cap);
}

// =============================================================================================
// shared helpers of the per-sample phases
// =============================================================================================
// pow(x, 0.25) of ComputeScore: the device evaluates the host libm's own algorithm (pow_quarter.h), bit for bit
PSC_DEV double pow_quarter(double x)
{
#ifdef PSC_CPU_SHIM
	return pow(x, 0.25); // the shim is compared with the oracle, which calls the host libm like the reference
#else
	return pow_quarter_libm(x);
#endif
}

PSC_DEV double compute_score(double cnt, double weight, double a, double A, double cor)
{ // TranscriptDecider.hpp:394-401
	if (a > A * 0.1) return (cnt * weight) * (1 + pow_quarter(a / A)) + cor;
	return (cnt * weight) * (1 + a / A) + cor;
}
PSC_DEV int tx_len(const CompView &v, const uint64_t *b, int first, int last)
{
	int len = 0;
	for (int q = first; q <= last; ++q)
		if (bt_test(b, q)) len += v.en[q] - v.st[q] + 1;
	return len;
}
PSC_DEV double to_fpkm(const SolveParams &p, double abundance, int len)
{ // ConvertTranscriptAbundanceToFPKM (TranscriptDecider.hpp:405-419), readCnt = 1e6
	double factor = p.matePaired ? 0.5 : 1;
	return abundance * factor / ((1000000 / 1000000.0) * (len / 1000.0));
}
PSC_DEV int len_from_fpkm(const SolveParams &p, double abundance, double fpkm)
{ // GetTranscriptLengthFromAbundanceAndFPKM (TranscriptDecider.hpp:421-427)
	double factor = p.matePaired ? 0.5 : 1;
	return int(abundance * factor / (fpkm / 1000.0) / (1000000 / 1000000.0) + 0.5);
}
PSC_DEV int gene_id_of(const CompView &v, const int *defGene, const uint64_t *b, int first, int last)
{ // GetTranscriptGeneId (TranscriptDecider.cpp:120-147)
	if (v.gene[first] != -2) return v.gene[first];
	if (v.gene[last] != -2) return v.gene[last];
	for (int q = first; q <= last; ++q)
		if (bt_test(b, q) && v.gene[q] != -2) return v.gene[q];
	int prev = -1;
	for (int q = first; q <= last; ++q) {
		if (!bt_test(b, q)) continue;
		if (prev >= 0 && v.en[prev] + 1 < v.st[q]) return defGene[(v.rs[prev] + 1) / 2];
		prev = q;
	}
	return defGene[0];
}

// candidate transcript accessors
struct CandView {
	int A, W;
	const uint64_t *bits;
	const int *first, *last;
	const int8_t *corr;
	int *sup1, *sup2;
};
PSC_DEV CandView cand_view(const SolveDev &sv, int c, int W)
{
	CandView k;
	long long r = sv.cand_row_off[c];
	k.A = sv.n_cand[c];
	k.W = W;
	k.bits = sv.cand_bits + sv.cand_bits_off[c];
	k.first = sv.cand_first + r;
	k.last = sv.cand_last + r;
	k.corr = sv.cand_corr + r;
	k.sup1 = sv.cand_sup1 + r;
	k.sup2 = sv.cand_sup2 + r;
	return k;
}

// pair j compatible with transcript (both mates): the btable predicate (.cpp:1435-1436)
PSC_DEV bool pair_in_tx(const TableView &t, int j, const uint64_t *tb, int tf, int tl, int W)
{
	int a = t.pi[j], b = t.pj[j];
	return compat(tb, tf, tl, false, t.bits + (long long)a * W, t.first[a], t.last[a], W) == 1 &&
	       compat(tb, tf, tl, false, t.bits + (long long)b * W, t.first[b], t.last[b], W) == 1;
}

// One row of the compatibility bit-matrix (btable[i], .cpp:1426-1442).  A mate pair is compatible when both
// of its constraints are, and a component has far fewer constraints than pairs: so the lanes first test
// the C constraints once (flags[]), then a pair costs two byte lookups; a warp ballot packs 32 pairs into
// one 32-bit half word.  Returns non-zero (in every lane) if any bit of the row is set.
PSC_DEV int build_pair_row(const Coop &co, const TableView &t, const uint64_t *tb, int tf, int tl, int W, int M, int MW,
                           uint64_t *row, uint8_t *flags)
{
	for (int c = co.lane; c < t.C; c += co.n)
		flags[c] = compat(tb, tf, tl, false, t.bits + (long long)c * W, t.first[c], t.last[c], W) == 1;
	co.sync();
	uint32_t *row32 = (uint32_t *)row;
	int any = 0;
	int n32 = 2 * MW;
	for (int k = 0; k < n32; ++k) {
		uint32_t mask = 0;
		int j0 = k * 32;
		if (co.n > 1) {
			int j = j0 + co.lane;
			bool in = j < M && flags[t.pi[j]] && flags[t.pj[j]];
			mask = co.ballot(in);
		} else {
			for (int b = 0; b < 32; ++b) {
				int j = j0 + b;
				if (j < M && flags[t.pi[j]] && flags[t.pj[j]]) mask |= 1u << b;
			}
		}
		if (co.lane == 0) row32[k] = mask;
		if (mask) any = 1;
	}
	co.sync(); // flags[] is rewritten for the next row
	return any;
}

// =============================================================================================
// PickTranscripts (TranscriptDecider.cpp:1368-2119): greedy quantitative set cover
// round2: candidate weights use the round-1 support (alltranscripts[i].FPKM); the compatibility
// matrix and the per-candidate rough abundances do not depend on the round, so the second round
// reuses what the first one left in the unit's scratch.
// output: pr.id / pr.abundance / pr.corr, returns the number picked
// =============================================================================================
// remember that candidate `cand` holds support for pair j (its csupport entry left zero): the redistribution
// step only has to look at these candidates.  Lists longer than PSC_OL fall back to the scan of all row owners.
PSC_DEV void pair_owner_add(const UnitScratch &sc, int j, int cand)
{
	int c = sc.ocnt[j];
	if (c < 0) return;
	int *L = sc.olist + (long long)j * PSC_OL;
	for (int q = 0; q < c; ++q)
		if (L[q] == cand) return;
	if (c == PSC_OL) {
		sc.ocnt[j] = -1;
		return;
	}
	int q = c;
	while (q > 0 && L[q - 1] > cand) {
		L[q] = L[q - 1];
		--q;
	}
	L[q] = cand;
	sc.ocnt[j] = c + 1;
}

PSC_DEV double pick_score(const UnitScratch &sc, const CandView &cd, const SolveParams &prm, int i, double cnt,
                          double maxAbundance, bool round2)
{
	if (cnt < 1) return -1; // "continue" of the reference: never selectable (all real scores are >= 1)
	cnt *= sc.covPortion[i];
	double weight = 1;
	if (round2 && cd.sup1[i] > 0 && prm.S > 1) weight = (1 + (double)cd.sup1[i] / prm.S);
	return compute_score(cnt, weight, sc.tAbund[i], maxAbundance, (double)cd.corr[i]);
}

PSC_DEV int pick_transcripts(const Coop &co, const CompView &v, const CandView &cd, const TableView &t,
                             const SolveParams &prm, UnitScratch &sc, int cap, bool round2, UnitPred &pr)
{
	int A = cd.A, M = t.M, W = v.W;
	if (M == 0 || A == 0) return 0;
	int MW = (M + 63) >> 6;
	double inf, maxAbundance;
	if (!round2) {
		// ---- inf ----
		inf = -1;
		for (int j = co.lane; j < M; j += co.n) {
			double x = t.support[t.pi[j]];
			if (t.support[t.pj[j]] > x) x = t.support[t.pj[j]];
			if (t.pnorm[j] > x) x = t.pnorm[j];
			if (x > inf) inf = x;
		}
		inf = co.max(inf);
		inf = inf + 1;
		// ---- compatibility matrix: warp ballot over 32 pairs at a time ----
		int anySet = 0;
		for (int i = 0; i < A; ++i)
			anySet |= build_pair_row(co, t, cd.bits + (long long)i * W, cd.first[i], cd.last[i], W, M, MW,
			                         sc.btable + (long long)i * MW, sc.flagsA);
		anySet = co.max(anySet);
		co.sync();
		if (co.lane == 0) {
			sc.persist[0] = inf;
			sc.persist[1] = anySet;
		}
		if (!anySet) {
			co.sync();
			return 0;
		}
		// ---- per-candidate rough abundance: lanes over candidates, sequential over pairs.
		// All pair abundances equal their normAbund here (tc[j].abundance = tc[j].normAbund) ----
		maxAbundance = -1;
		int *reach = sc.reach + (long long)co.lane * v.n;
		for (int i = co.lane; i < A; i += co.n) {
			const uint64_t *tb = cd.bits + (long long)i * W;
			const uint64_t *bt = sc.btable + (long long)i * MW;
			int tf = cd.first[i], tl = cd.last[i];
			double value = inf;
			int tag = -1;
			int seIdxCnt = 0;
			for (int w = 0; w < W; ++w) seIdxCnt += popc64(tb[w]);
			int compatibleCnt = 0;
			double avg = 0;
			for (int q = tf; q <= tl; ++q) reach[q] = -1;
			for (int w = 0; w < MW; ++w) {
				uint64_t word = bt[w];
				while (word) {
					int j = w * 64 + ctz64(word);
					word &= word - 1;
					double pa = t.pnorm[j];
					if (!(pa > 0)) continue;
					++compatibleCnt;
					double adjust = pa;
					int ci = t.pi[j], cj = t.pj[j];
					int fi = t.first[ci], li = t.last[ci], fj = t.first[cj], lj = t.last[cj];
					if (seIdxCnt > 1) {
						int si = fi + li, sj = fj + lj;
						bool iEdge = si == 2 * tf || si == 2 * tl;
						bool jEdge = sj == 2 * tf || sj == 2 * tl;
						if (ci == cj && iEdge) adjust = inf;
						else if (ci != cj && iEdge) adjust = t.norm[cj];
						else if (ci != cj && jEdge) adjust = t.norm[ci];
					}
					if (adjust < value) { value = adjust; tag = j; }
					avg += pa;
					// a compatible constraint repeats the transcript's bits on [first, last]: it covers
					// every chain (consecutive subexon pair of the transcript) inside that range
					if (li > reach[fi]) reach[fi] = li;
					if (lj > reach[fj]) reach[fj] = lj;
				}
			}
			// covered chains (.cpp:1607-1629, 1660-1673)
			int cnt = 0;
			bool uncovered = false;
			{
				int prev = -1, far = -1;
				for (int q = tf; q <= tl; ++q) {
					if (!bt_test(tb, q)) continue;
					if (prev >= 0) {
						if (far >= q) ++cnt;
						else uncovered = true;
					}
					if (reach[q] > far) far = reach[q];
					prev = q;
				}
			}
			// together penalty (.cpp:1634-1656): regions of >= 2 IR/overhang subexons partially used
			for (int r0 = 0; r0 < v.n;) {
				int r1;
				for (r1 = r0 + 1; r1 < v.n; ++r1)
					if (v.st[r1] > v.en[r1 - 1] + 1) break;
				int tot = 0, got = 0;
				for (int q = r0; q < r1; ++q) {
					bool tg = (v.lt[q] == 2 && v.rt[q] == 1) || (v.lt[q] == 0 && v.rt[q] == 1) || (v.lt[q] == 2 && v.rt[q] == 0);
					if (tg) {
						++tot;
						if (bt_test(tb, q)) ++got;
					}
				}
				if (tot > 1 && got > 0 && got < tot) value /= 2;
				r0 = r1;
			}
			if (uncovered) value = 0;
			double portion;
			if (seIdxCnt > 1) portion = (double)cnt / (double)(seIdxCnt - 1);
			else portion = 1;
			if (portion == 0) portion = (double)0.5 / (seIdxCnt);
			sc.covPortion[i] = portion;
			if (tag == -1) value = 0;
			if (value > maxAbundance) maxAbundance = value;
			sc.tAbund[i] = value;
			if (tag != -1) avg /= compatibleCnt;
			sc.avgAbund[i] = avg;
		}
		maxAbundance = co.max(maxAbundance);
		co.sync();
		if (maxAbundance == 0) {
			for (int i = co.lane; i < A; i += co.n) sc.tAbund[i] = sc.covPortion[i];
			maxAbundance = 1;
		}
		if (co.lane == 0) sc.persist[2] = maxAbundance;
		co.sync();
	} else {
		inf = sc.persist[0];
		if (sc.persist[1] == 0) return 0;
		maxAbundance = sc.persist[2];
	}
	// ---- per-round state ----
	for (int j = co.lane; j < M; j += co.n) {
		sc.pairAbund[j] = t.pnorm[j];
		sc.ocnt[j] = 0;
	}
	co.sync();
	for (int i = co.lane; i < A; i += co.n) {
		sc.abund[i] = 0;
		sc.slot[i] = -1;
		const uint64_t *bt = sc.btable + (long long)i * MW;
		double cnt = 0;
		for (int w = 0; w < MW; ++w) {
			uint64_t word = bt[w];
			while (word) {
				int j = w * 64 + ctz64(word);
				word &= word - 1;
				if (sc.pairAbund[j] > 0) cnt += t.pec[j];
			}
		}
		sc.coverCnt[i] = cnt;
		sc.score[i] = pick_score(sc, cd, prm, i, cnt, maxAbundance, round2);
	}
	co.sync();
	// ---- quantitative set cover (.cpp:1827-2068) ----
	int poolUsed = 0;
	for (;;) {
		// arg-max: lexicographic (score, avgAbund) maximum, lowest index on full ties
		double bestScore = -1, bestAvg = 0;
		int bestI = -1;
		for (int i = co.lane; i < A; i += co.n) {
			double s = sc.score[i];
			if (s < 0) continue;
			if (bestI == -1 || s > bestScore || (s == bestScore && bestAvg < sc.avgAbund[i])) {
				bestScore = s;
				bestAvg = sc.avgAbund[i];
				bestI = i;
			}
		}
		if (co.n > 1) {
			for (int o = 16; o > 0; o >>= 1) {
				double os = co.bcast(bestScore, co.lane ^ o), oa = co.bcast(bestAvg, co.lane ^ o);
				int oi = co.bcast(bestI, co.lane ^ o);
				bool take = false;
				if (oi >= 0) {
					if (bestI < 0) take = true;
					else if (os > bestScore) take = true;
					else if (os == bestScore) {
						if (oa > bestAvg) take = true;
						else if (oa == bestAvg && oi < bestI) take = true;
					}
				}
				if (take) { bestScore = os; bestAvg = oa; bestI = oi; }
			}
		}
		int maxtag = bestI;
		if (maxtag == -1) break;
		const uint64_t *btm = sc.btable + (long long)maxtag * MW;
		// pair to deplete: smallest remaining abundance, last index on ties (.cpp:1913-1923)
		double update = inf;
		int updateTag = 0;
		{
			double lu = inf;
			int lt = -1;
			for (int j = co.lane; j < M; j += co.n) {
				if (!bt_test(btm, j)) continue;
				double pa = sc.pairAbund[j];
				if (pa > 0 && pa <= lu) { lu = pa; lt = j; }
			}
			if (co.n > 1) {
				for (int o = 16; o > 0; o >>= 1) {
					double ou = co.bcast(lu, co.lane ^ o);
					int ot = co.bcast(lt, co.lane ^ o);
					if (ot >= 0 && (lt < 0 || ou < lu || (ou == lu && ot > lt))) { lu = ou; lt = ot; }
				}
			}
			if (lt >= 0) { update = lu; updateTag = lt; }
		}
		// the chosen candidate needs a csupport row
		if (sc.slot[maxtag] < 0) {
			if (poolUsed >= cap) return -1; // cannot happen: one new candidate per iteration, <= M+1 iterations
			double *row = sc.pool + (long long)poolUsed * M;
			for (int j = co.lane; j < M; j += co.n) row[j] = 0;
			co.sync();
			if (co.lane == 0) {
				sc.slot[maxtag] = poolUsed;
				int q = poolUsed; // keep the owner list sorted by candidate index
				while (q > 0 && sc.picked[q - 1] > maxtag) { sc.picked[q] = sc.picked[q - 1]; --q; }
				sc.picked[q] = maxtag;
			}
			++poolUsed;
			co.sync();
		}
		double *rowm = sc.pool + (long long)sc.slot[maxtag] * M;
		int supportCnt = 0, deplCnt = 0;
		int *dcount = sc.depl + M;
		if (co.lane == 0) *dcount = 0;
		co.sync();
		// phase A, lanes over the words of maxtag's row: pairs that still have abundance (.cpp:1957-1975).
		// Their contribution to alltranscripts[maxtag].abundance is parked in contrib[] and added in
		// index order by phase B, so the fp64 sum keeps the reference's order.
		{
			int ls = 0;
			for (int j = co.lane; j < M; j += co.n) {
				if (!bt_test(btm, j)) continue;
				double pa = sc.pairAbund[j];
				if (pa > 0) {
					pa -= 1 * update;
					double factor = t.pec[j];
					double tmp = (t.psup[j] * update / t.pnorm[j] * factor);
					if (rowm[j] == 0 && tmp != 0) pair_owner_add(sc, j, maxtag);
					rowm[j] += tmp;
					sc.contrib[j] = tmp;
					if (pa <= 0) sc.depl[atomic_add(dcount, 1)] = j;
					if (pa < 0) pa = 0;
					sc.pairAbund[j] = pa;
					++ls;
				} else
					sc.contrib[j] = -1; // depleted before this iteration: phase B looks at it
			}
			supportCnt = co.sum(ls);
		}
		co.sync();
		// phase B, in pair order with all lanes in lockstep: accumulate, and redistribute support for
		// pairs that were already depleted and that this candidate never got a share of (.cpp:1976-2054).
		// Pairs are visited sequentially (whether an owner still has abundance depends on the pairs
		// before); inside one pair the lanes take one row owner each and the fp64 take-out sum is
		// folded in owner order.
		{
			double abm = sc.abund[maxtag];
			const double tam = sc.tAbund[maxtag];
			for (int w = 0; w < MW; ++w) {
				uint64_t word = btm[w];
				while (word) {
					int j = w * 64 + ctz64(word);
					word &= word - 1;
					double cj = sc.contrib[j];
					if (cj >= 0) {
						abm += cj;
					} else if (rowm[j] == 0) {
						// rowm[j] == 0 here, so (csupport[i][j] + rowm[j]) * a / (a + b) - rowm[j] is csupport * a / (a + b).
						// Only candidates that were picked before can have abundance > 0: walk the (index-sorted)
						// list of row owners instead of all A candidates.
						double takeOut = 0, factor = t.pec[j];
						// the candidates that ever held support for pair j (ascending), or all row owners
						const int oc = sc.ocnt[j];
						const int *ol = oc >= 0 ? sc.olist + (long long)j * PSC_OL : sc.picked;
						const int on = oc >= 0 ? oc : poolUsed;
						// first round: lane q looks at list entry q and keeps what it found for the update below
						double tmp1 = 0;
						int i1 = -1;
						if (co.lane < on) {
							int i = ol[co.lane];
							if (i != maxtag && sc.abund[i] > 0 && bt_test(sc.btable + (long long)i * MW, j)) {
								i1 = i;
								tmp1 = sc.pool[(long long)sc.slot[i] * M + j] * tam / (tam + sc.tAbund[i]);
							}
						}
						{
							uint32_t m = co.ballot(tmp1 > 0);
							while (m) {
								int bsrc = ctz32(m);
								m &= m - 1;
								takeOut += co.bcast(tmp1, bsrc);
							}
						}
						for (int p0 = co.n; p0 < on; p0 += co.n) { // only the fallback list can be this long
							int pq = p0 + co.lane;
							double tmp = 0;
							if (pq < on) {
								int i = ol[pq];
								if (i != maxtag && sc.abund[i] > 0 && bt_test(sc.btable + (long long)i * MW, j))
									tmp = sc.pool[(long long)sc.slot[i] * M + j] * tam / (tam + sc.tAbund[i]);
							}
							uint32_t m = co.ballot(tmp > 0);
							while (m) {
								int bsrc = ctz32(m);
								m &= m - 1;
								takeOut += co.bcast(tmp, bsrc);
							}
						}
						// nothing to take from: the share stays 0 and the sums below would add 0.0
						if (takeOut > 0) {
							double ratio = 1, takeOutFactor = 0.5;
							if (update < t.pnorm[j]) {
								if (takeOut > (t.psup[j] * update / t.pnorm[j] * factor) * takeOutFactor)
									ratio = (t.psup[j] * update / t.pnorm[j] * factor) * takeOutFactor / takeOut;
							} else {
								if (takeOut > (t.psup[j] * factor) * takeOutFactor)
									ratio = (t.psup[j] * factor) * takeOutFactor / takeOut;
							}
							if (tmp1 > 0) { // same value the first round computed: nothing it depends on has changed
								double tmp = tmp1 * ratio;
								sc.pool[(long long)sc.slot[i1] * M + j] -= tmp;
								sc.abund[i1] -= tmp;
							}
							for (int pq = co.n + co.lane; pq < on; pq += co.n) {
								int i = ol[pq];
								if (i != maxtag && sc.abund[i] > 0 && bt_test(sc.btable + (long long)i * MW, j)) {
									double *ri = sc.pool + (long long)sc.slot[i] * M;
									double tmp = ri[j] * tam / (tam + sc.tAbund[i]);
									if (tmp > 0) {
										tmp *= ratio;
										ri[j] -= tmp;
										sc.abund[i] -= tmp;
									}
								}
							}
							co.sync(); // the next pair may hand the same candidate to another lane
							double tmp = takeOut * ratio;
							if (co.lane == 0) {
								if (tmp != 0) pair_owner_add(sc, j, maxtag);
								rowm[j] += tmp;
							}
							abm += tmp;
						}
					}
				}
			}
			if (co.lane == 0) {
				sc.abund[maxtag] = abm;
				// tc[updateTag].abundance = 0 (.cpp:2063); when it is not one of maxtag's pairs (no pair was
				// left: update == inf) the reference zeroes pair 0 without touching coverCnt and stops.
				sc.pairAbund[updateTag] = 0;
			}
		}
		co.sync();
		deplCnt = *dcount;
		if (supportCnt == 0) break;
		// coverCnt bookkeeping (.cpp:1965-1973): every candidate sharing a pair that just ran dry
		if (deplCnt > 0)
			for (int i = co.lane; i < A; i += co.n) {
				const uint64_t *bt = sc.btable + (long long)i * MW;
				double cc = sc.coverCnt[i];
				bool changed = false;
				for (int q = 0; q < deplCnt; ++q)
					if (bt_test(bt, sc.depl[q])) { cc -= t.pec[sc.depl[q]]; changed = true; }
				if (changed) {
					sc.coverCnt[i] = cc;
					sc.score[i] = pick_score(sc, cd, prm, i, cc, maxAbundance, round2);
				}
			}
		co.sync();
	}
	// ---- collect (.cpp:2070-2085) ----
	int T = 0;
	for (int b0 = 0; b0 < A; b0 += co.n) {
		int i = b0 + co.lane;
		bool in = i < A && sc.abund[i] > 0;
		uint32_t m = co.ballot(in);
		if (in) {
			int p = T + co.rank_in(m);
			pr.id[p] = i;
			pr.abundance[p] = sc.abund[i];
			pr.corr[p] = sc.tAbund[i] == 0 ? -1 : 0;
		}
		T += popc32(m);
	}
	co.sync();
	return T;
}

} // namespace pscdev

#include "k_solve2.cuh"
#include "k_dp.cuh"
