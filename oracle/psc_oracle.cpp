// psc_oracle.cpp -- CPU restatement of PsiCLASS stage 3.  TEST INFRASTRUCTURE (see psc_oracle.h).
// Citations are to /root/reference (PsiCLASS v1.0.3).  The quirks Q1..Q14 of SURVEY.md §7 are
// reproduced on purpose; each is marked where it occurs.
#include "psc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

namespace orc {

// ------------------------------------------------------------------------------------------
// bit vectors (BitTable.hpp:20-468; bit i <-> subexon i, LSB first)
// ------------------------------------------------------------------------------------------
static inline Bits bits_new(int n) { return Bits((size_t)nwords(n), 0); }
static inline void bits_reset(Bits &b) { std::fill(b.begin(), b.end(), 0); }
static inline bool bits_equal(const Bits &a, const Bits &b) { return a == b; }
static inline bool bits_zero(const Bits &a)
{
	for (uint64_t w : a)
		if (w) return false;
	return true;
}
static inline int bits_count(const Bits &a)
{
	int c = 0;
	for (uint64_t w : a) c += __builtin_popcountll(w);
	return c;
}
static inline void bits_ones(const Bits &a, std::vector<int> &out)
{
	out.clear();
	for (size_t w = 0; w < a.size(); ++w) {
		uint64_t x = a[w];
		while (x) {
			out.push_back((int)(w * 64 + __builtin_ctzll(x)));
			x &= x - 1;
		}
	}
}
// BitTable::MaskRegionOutside(s, e): keep bits [s, e] (BitTable.hpp:182-213)
static inline void bits_keep(Bits &a, int s, int e)
{
	for (size_t w = 0; w < a.size(); ++w) {
		int lo = (int)w * 64, hi = lo + 63;
		if (hi < s || lo > e) {
			a[w] = 0;
			continue;
		}
		uint64_t m = ~(uint64_t)0;
		if (s > lo) m &= ~(uint64_t)0 << (s - lo);
		if (e < hi) m &= ~(uint64_t)0 >> (hi - e);
		a[w] &= m;
	}
}
// BitTable::GetFirstDifference (BitTable.hpp:373-410)
static inline int bits_first_diff(const Bits &a, const Bits &b)
{
	for (size_t w = 0; w < a.size(); ++w)
		if (a[w] != b[w])
			return (int)(w * 64 + __builtin_ctzll(a[w] ^ b[w]));
	return -1;
}

// ------------------------------------------------------------------------------------------
// Constraints.cpp
// ------------------------------------------------------------------------------------------

// Constraints::ConvertAlignmentToBitTable (Constraints.cpp:4-72)
static bool alignment_to_bits(const int32_t *sa, const int32_t *sb, int segCnt, const std::vector<Subexon> &se,
                              int seStart, Constraint &ct)
{
	int seCnt = (int)se.size();
	int k = seStart;
	ct.v = bits_new(seCnt);
	for (int i = 0; i < segCnt; ++i) {
		int leftIdx = -1, rightIdx = -1;
		for (; k < seCnt; ++k) {
			if (se[k].start > sb[i])
				break;
			if (sa[i] > se[k].end)
				continue;
			if ((se[k].start >= sa[i] && se[k].end <= sb[i]) ||
			    (i == 0 && se[k].start < sa[i] && se[k].end <= sb[i]) ||
			    (i == segCnt - 1 && se[k].start >= sa[i] && se[k].end > sb[i]) ||
			    (i == 0 && i == segCnt - 1 && se[k].start < sa[i] && se[k].end > sb[i])) {
				if (leftIdx == -1)
					leftIdx = k;
				rightIdx = k;
				bit_set(ct.v, k);
			} else
				return false;
		}
		if (leftIdx == -1)
			return false;
		if (!((se[leftIdx].lt == 0 || se[leftIdx].start <= sa[i]) &&
		      (se[rightIdx].rt == 0 || se[rightIdx].end >= sb[i])))
			return false;
		if (i > 0) {
			const std::vector<int> &nx = se[ct.last].next;
			if (std::find(nx.begin(), nx.end(), leftIdx) == nx.end())
				return false;
		}
		for (int j = leftIdx + 1; j <= rightIdx; ++j)
			if (se[j].start > se[j - 1].end + 1)
				return false;
		if (i == 0)
			ct.first = leftIdx;
		ct.last = rightIdx;
	}
	return true;
}

// Constraints::CompSortConstraints (Constraints.hpp:248-264)
static bool constraint_less(const Constraint &a, const Constraint &b)
{
	if (a.first != b.first)
		return a.first < b.first;
	int d = bits_first_diff(a.v, b.v);
	if (d == -1)
		return false;
	return !bit_test(a.v, d);
}

// Constraints::CoalesceSameConstraints (Constraints.cpp:74-150).  std::sort leaves the order of
// equal vectors unspecified, and the weights 10/NH of equal vectors are summed in that order (Q9:
// inexact for NH = 3).  The same std::sort call on the same sequence with the same comparator makes
// the same moves whatever the element type, so calling it here reproduces the reference's order as
// long as both are built against the same libstdc++ (checked against oracle/_ref/ref_dump).
static void coalesce(ConstraintSet &cs, std::vector<int> *newIdxOut = nullptr)
{
	int size = (int)cs.c.size();
	for (int i = 0; i < size; ++i)
		cs.c[i].info = i;
	std::vector<int> newIdx((size_t)size, 0);
	if (size > 0) {
		std::sort(cs.c.begin(), cs.c.end(), constraint_less);
		int k = 0;
		newIdx[cs.c[0].info] = 0;
		for (int i = 1; i < size; ++i) {
			if (bits_equal(cs.c[k].v, cs.c[i].v)) {
				cs.c[k].weight += cs.c[i].weight;
				cs.c[k].support += cs.c[i].support;
				cs.c[k].uniq += cs.c[i].uniq;
				cs.c[k].maxReadLen = std::max(cs.c[k].maxReadLen, cs.c[i].maxReadLen);
			} else {
				++k;
				if (k != i)
					cs.c[k] = cs.c[i];
			}
			newIdx[cs.c[i].info] = k;
		}
		cs.c.resize((size_t)k + 1);
	}
	size = (int)cs.p.size();
	if (size > 0) {
		for (auto &m : cs.p) {
			m.i = newIdx[m.i];
			m.j = newIdx[m.j];
		}
		std::stable_sort(cs.p.begin(), cs.p.end(), [](const MatePair &a, const MatePair &b) {
			return a.i != b.i ? a.i < b.i : a.j < b.j;
		});
		int k = 0;
		for (int i = 1; i < size; ++i) {
			if (cs.p[i].i == cs.p[k].i && cs.p[i].j == cs.p[k].j) {
				cs.p[k].support += cs.p[i].support;
				cs.p[k].uniq += cs.p[i].uniq;
			} else {
				++k;
				cs.p[k] = cs.p[i];
			}
		}
		cs.p.resize((size_t)k + 1);
	}
	if (newIdxOut) newIdxOut->swap(newIdx);
}

// Constraints::ComputeNormAbund (Constraints.cpp:152-238)
static void norm_abund(const std::vector<Subexon> &se, ConstraintSet &cs, int sampleReadLen)
{
	std::vector<int> ind;
	for (auto &c : cs.c) {
		int readLen = c.maxReadLen;
		if (c.first + 1 < c.last) {
			bits_ones(c.v, ind);
			for (size_t j = 1; j + 1 < ind.size(); ++j)
				readLen -= se[ind[j]].end - se[ind[j]].start + 1;
		}
		int eff;
		if (c.first == c.last) {
			eff = (se[c.first].end - readLen + 1) - se[c.first].start + 1;
			if (eff <= 0)
				eff = (se[c.first].end - se[c.first].start + 1) / 2 + 1;
		} else {
			int a = c.first, b = c.last, start, end;
			start = se[a].end + 1 - (readLen - 1);
			if (start < se[a].start)
				start = se[a].start;
			if (se[b].end - se[b].start + 1 >= readLen - 1 || se[b].rt == 0)
				end = se[a].end;
			else
				end = se[a].end + 1 - (readLen - (se[b].end - se[b].start + 1));
			if (end < start)
				end = se[a].start;
			eff = end - start + 1;
		}
		c.normAbund = (double)c.weight / (double)eff;
		if ((se[c.first].lt == 0 && se[c.first].end - se[c.first].start + 1 >= 8 * sampleReadLen) ||
		    (se[c.last].rt == 0 && se[c.last].end - se[c.last].start + 1 >= 8 * sampleReadLen))
			c.normAbund *= 2;
		c.abundance = c.normAbund;
	}
	for (auto &m : cs.p) {
		double a = cs.c[m.i].normAbund, b = cs.c[m.j].normAbund;
		m.normAbund = a < b ? a : b;
		if (m.i != m.j) {
			if (se[cs.c[m.i].first].lt == 0 && cs.c[m.i].first == cs.c[m.i].last && a < b)
				m.normAbund = b;
			else if (se[cs.c[m.j].last].rt == 0 && cs.c[m.j].first == cs.c[m.j].last && a > b)
				m.normAbund = a;
		}
		m.abundance = m.normAbund;
	}
}

static int chr_of(const Reads &r, int64_t idx)
{
	// chr_off is small; linear scan from the back is fine for test infrastructure
	int c = (int)r.chr_off.size() - 2;
	while (c > 0 && r.chr_off[c] > idx)
		--c;
	return c;
}

// Constraints::BuildConstraints (Constraints.cpp:240-438) + MateReadIds (Constraints.hpp:52-231)
void build_constraints(const Component &comp, Reads &r, bool primaryAsUnique, ConstraintSet &out)
{
	const std::vector<Subexon> &se = comp.se;
	int seCnt = (int)se.size();
	std::vector<Constraint>().swap(out.c); // capacity restarts at 0 for every component (Constraints.cpp:247-254)
	out.p.clear();
	int tag = 0;
	int coalesceThreshold = 16384;
	// mate bookkeeping: entries keyed by matePos; cache = ids whose stored matePos == cachedMatePos.
	std::multimap<int, std::pair<uint64_t, int>> pending;
	std::map<uint64_t, int> cache;
	int cachedMatePos = -1;

	for (;;) {
		if (!r.held) {
			if (r.cursor >= r.n) {
				r.last_read = r.n; // EOF: samread() failed (alignments.hpp:164-168)
				break;
			}
			r.last_read = r.cursor;
		}
		r.held = false;
		int64_t x = r.cursor;
		int chr = chr_of(r, x);
		if (chr < comp.chr) {
			++r.cursor;
			continue;
		}
		if (chr > comp.chr) {
			r.held = true;
			break;
		}
		const int32_t *sa = r.seg_a + r.seg_off[x], *sb = r.seg_b + r.seg_off[x];
		int segCnt = (int)(r.seg_off[x + 1] - r.seg_off[x]);
		for (; tag < seCnt && se[tag].end < sa[0]; ++tag)
			;
		if (tag >= seCnt) {
			r.held = true;
			break;
		}
		++r.cursor;
		if (sb[segCnt - 1] < se[tag].start)
			continue;
		uint8_t fl = r.flags[x];
		bool primary = fl & 1;
		int uniq = primaryAsUnique ? (primary ? 1 : 0) : ((fl & 16) ? 1 : 0);
		Constraint ct;
		ct.weight = 1.0 / r.nh[x];
		ct.weight *= 10; // Q1: IsGCRich() is "always true" in the -O3 reference build
		ct.normAbund = 0;
		ct.support = 1;
		ct.uniq = uniq;
		int cover = 0;
		for (int i = 0; i < segCnt; ++i)
			cover += sb[i] - sa[i] + 1;
		ct.maxReadLen = cover;
		if (!(primary && alignment_to_bits(sa, sb, segCnt, se, tag, ct)))
			continue;
		bool validClip = true;
		if (fl & 2) {
			if ((ct.first < seCnt - 1 && se[ct.first].end + 1 == se[ct.first + 1].start) ||
			    !se[ct.first].prev.empty() || sb[0] - sa[0] + 1 <= cover / 3.0)
				validClip = false;
		}
		if (fl & 4) {
			int t = segCnt - 1;
			if ((ct.last > 0 && se[ct.last].start - 1 == se[ct.last - 1].end) || !se[ct.last].next.empty() ||
			    sb[t] - sa[t] + 1 <= cover / 3.0)
				validClip = false;
		}
		if (!validClip)
			continue;
		out.c.push_back(ct);
		int me = (int)out.c.size() - 1;
		if (fl & 32) { // mate on the same chromosome (Constraints.cpp:337)
			int matePos = r.mate_pos[x], pos = sa[0];
			if (matePos < pos) {
				// MateReadIds::Query(id, pos)
				if (pos > cachedMatePos) {
					cache.clear();
					pending.erase(pending.begin(), pending.lower_bound(pos));
					auto range = pending.equal_range(pos);
					for (auto it = range.first; it != range.second; ++it)
						cache[it->second.first] = it->second.second; // later insertions win
					pending.erase(range.first, range.second);
					cachedMatePos = pos;
				}
				auto hit = cache.find(r.name_key[x]);
				if (hit != cache.end()) {
					MatePair nm;
					nm.i = hit->second;
					nm.j = me;
					nm.support = 1;
					nm.uniq = uniq;
					nm.effectiveCount = 2;
					out.p.push_back(nm);
				}
			} else if (matePos > pos) {
				pending.insert({matePos, {r.name_key[x], me}});
			} else if (fl & 8) { // same coordinate: only the first mate makes the (k,k) pair
				MatePair nm;
				nm.i = nm.j = me;
				nm.support = 1;
				nm.uniq = uniq;
				nm.effectiveCount = 2;
				out.p.push_back(nm);
			}
		}
		// capacity-triggered partial coalescing (Constraints.cpp:378-390).  The reference tests this after every
		// primary convertible alignment, pushed or not; right after a coalesce size > threshold cannot hold
		// together with size == capacity, so testing after pushes only is equivalent.
		if ((int)out.c.size() > coalesceThreshold && out.c.size() == out.c.capacity()) {
			std::vector<int> newIdx;
			coalesce(out, &newIdx);
			for (auto &e : pending) e.second.second = newIdx[(size_t)e.second.second]; // MateReadIds::UpdateIdx
			for (auto &e : cache) e.second = newIdx[(size_t)e.second];
			if (out.c.size() >= out.c.capacity() / 2) coalesceThreshold *= 2;
		}
	}
	coalesce(out);
	if (r.frag_stdev == 0) { // single-end library (Constraints.cpp:405-422)
		out.p.clear();
		for (int i = 0; i < (int)out.c.size(); ++i) {
			MatePair nm;
			nm.i = nm.j = i;
			nm.support = out.c[i].support;
			nm.uniq = out.c[i].uniq;
			nm.effectiveCount = 1;
			out.p.push_back(nm);
		}
	}
	norm_abund(se, out, r.read_len);
}

// ------------------------------------------------------------------------------------------
// TranscriptDecider.cpp
// ------------------------------------------------------------------------------------------
namespace {

struct Transcript {
	Bits v;
	double abundance = 0, corr = 0, FPKM = 0;
	std::vector<double> csupport;
	int first = 0, last = 0;
	bool partial = false;
	int id = 0;
};

struct Dp {
	Bits v;
	int first = -1, last = -1, cnt = -1;
	double cover = -1;
	double minAbundance = -1;
	int timeStamp = -1;
	int strand = 0;
	int epoch = -1;
};

inline bool same_strand(int a, int b) { return a == 0 || b == 0 || a == b; }

struct Decider {
	const Params &prm;
	std::vector<Subexon> &se;
	int seCnt;
	int baseGeneId = 0, usedGeneId = 0, defaultGeneId[2] = {-1, -1};
	int hashMax = 100003;

	Decider(const Params &p, std::vector<Subexon> &s) : prm(p), se(s), seCnt((int)s.size()) {}

	// TranscriptDecider::IsConstraintInTranscript (TranscriptDecider.cpp:181-233)
	int compat(const Bits &tv, int tfirst, int tlast, bool tpartial, const Bits &cv, int cfirst, int clast) const
	{
		if (cfirst < tfirst || cfirst > tlast || !bit_test(tv, cfirst) || (!tpartial && !bit_test(tv, clast)))
			return 0;
		int s = cfirst, e = clast;
		bool retPartial = false;
		if (e > tlast) {
			if (tpartial) {
				e = tlast;
				retPartial = true;
			} else
				return 0;
		}
		size_t W = tv.size();
		for (size_t w = 0; w < W; ++w) {
			int lo = (int)w * 64, hi = lo + 63;
			uint64_t m = 0;
			if (!(hi < s || lo > e)) {
				m = ~(uint64_t)0;
				if (s > lo) m &= ~(uint64_t)0 << (s - lo);
				if (e < hi) m &= ~(uint64_t)0 >> (hi - e);
			}
			uint64_t t = tv[w] & m;
			uint64_t c = cv[w];
			if (clast > tlast) { // keep [0, e] of the constraint
				uint64_t mc = 0;
				if (lo <= e) {
					mc = ~(uint64_t)0;
					if (e < hi) mc &= ~(uint64_t)0 >> (hi - e);
				}
				c &= mc;
			}
			if (t != c)
				return 0;
		}
		return retPartial ? 2 : 1;
	}
	int compat(const Transcript &t, const Constraint &c) const
	{
		return compat(t.v, t.first, t.last, t.partial, c.v, c.first, c.last);
	}

	// ComputeScore (TranscriptDecider.hpp:394-401)
	static double score(double cnt, double weight, double a, double A, double cor)
	{
		if (a > A * 0.1)
			return (cnt * weight) * (1 + pow(a / A, 0.25)) + cor;
		return (cnt * weight) * (1 + a / A) + cor;
	}

	int txLen(const Bits &v) const
	{
		int len = 0;
		std::vector<int> ind;
		bits_ones(v, ind);
		for (int i : ind) len += se[i].end - se[i].start + 1;
		return len;
	}
	// ConvertTranscriptAbundanceToFPKM (TranscriptDecider.hpp:405-419), readCnt = 1e6
	void toFPKM(Transcript &t) const
	{
		int len = txLen(t.v);
		double factor = prm.matePaired ? 0.5 : 1;
		t.FPKM = t.abundance * factor / ((1000000 / 1000000.0) * (len / 1000.0));
	}
	// GetTranscriptLengthFromAbundanceAndFPKM (TranscriptDecider.hpp:421-427)
	int lenFromFPKM(double abundance, double FPKM) const
	{
		double factor = prm.matePaired ? 0.5 : 1;
		return int(abundance * factor / (FPKM / 1000.0) / (1000000 / 1000000.0) + 0.5);
	}

	// GetTranscriptGeneId (TranscriptDecider.cpp:120-147)
	int geneIdOf(const std::vector<int> &ind) const
	{
		int size = (int)ind.size();
		for (int i = 0; i < size; ++i)
			if (se[ind[i]].geneId != -2)
				return se[ind[i]].geneId;
		for (int i = 0; i < size - 1; ++i)
			if (se[ind[i]].end + 1 < se[ind[i + 1]].start)
				return defaultGeneId[(se[ind[i]].rs + 1) / 2];
		return defaultGeneId[0];
	}
	int geneIdOf(const Transcript &t) const
	{
		if (se[t.first].geneId != -2) return se[t.first].geneId;
		if (se[t.last].geneId != -2) return se[t.last].geneId;
		std::vector<int> ind;
		bits_ones(t.v, ind);
		return geneIdOf(ind);
	}

	// SubTranscriptCount (TranscriptDecider.cpp:284-301); int arithmetic wraps (Q4)
	int subCount(int tag, std::vector<int> &f)
	{
		if (f[tag] != -1)
			return f[tag];
		int ret = 0;
		if (se[tag].canEnd)
			ret = 1;
		for (int nx : se[tag].next)
			ret = (int)((uint32_t)ret + (uint32_t)subCount(nx, f));
		if (ret == 0)
			ret = 1;
		return f[tag] = ret;
	}

	// CompSortTranscripts (TranscriptDecider.hpp:360-374) + CoalesceSameTranscripts (.cpp:303-328)
	static bool tx_less(const Transcript &a, const Transcript &b)
	{
		if (a.first != b.first) return a.first < b.first;
		int d = bits_first_diff(a.v, b.v);
		if (d == -1) return false;
		return bit_test(a.v, d);
	}
	static void coalesceTx(std::vector<Transcript> &t)
	{
		if (t.empty()) return;
		std::stable_sort(t.begin(), t.end(), tx_less);
		size_t k = 0;
		for (size_t i = 1; i < t.size(); ++i) {
			if (bits_equal(t[k].v, t[i].v))
				t[k].abundance += t[i].abundance;
			else {
				++k;
				if (i != k) t[k] = t[i];
			}
		}
		t.resize(k + 1);
	}

	// EnumerateTranscript (TranscriptDecider.cpp:330-372); correlation.Query()==0 without --ls
	void enumerate(int tag, int strand, std::vector<int> &visit, int vcnt, double corScore,
	               std::vector<Transcript> &all)
	{
		visit[vcnt] = tag;
		double minCor = corScore;
		for (int i = 0; i < vcnt; ++i)
			if (0.0 < minCor) minCor = 0.0;
		if (se[tag].canEnd) {
			Transcript t;
			t.v = bits_new(seCnt);
			for (int i = 0; i <= vcnt; ++i) bit_set(t.v, visit[i]);
			t.first = visit[0];
			t.last = visit[vcnt];
			t.partial = false;
			t.corr = minCor;
			all.push_back(t);
		}
		for (int a : se[tag].next) {
			if (!same_strand(se[tag].rs, strand) && se[a].start > se[tag].end + 1)
				continue;
			int backup = strand;
			if (se[a].start > se[tag].end + 1 && strand == 0)
				strand = se[tag].rs;
			enumerate(a, strand, visit, vcnt + 1, minCor, all);
			strand = backup;
		}
	}

	// ---------------- DP (TranscriptDecider.cpp:374-1201) ----------------
	struct DpAttr {
		std::vector<Dp> f1;
		std::vector<Dp> f2; // seCnt*seCnt when seCnt <= 10000
		bool hasF2 = false;
		std::vector<Dp> hash;
		bool forAbundance = false;
		std::set<long long> uncovered;
		double minAbundance = 0;
		int timeStamp = 0;
		int epoch = 0; // replaces the per-call ResetDpContent sweep (.cpp:882-888)
	};

	Dp &memoSlot(Dp &d, DpAttr &attr)
	{
		if (d.epoch != attr.epoch) { // behaves like a freshly Reset entry
			d = Dp();
			d.epoch = attr.epoch;
		}
		return d;
	}

	void setDp(Dp &a, const Dp &b, const DpAttr &attr)
	{
		a.v = b.v;
		a.first = b.first;
		a.last = b.last;
		a.cnt = b.cnt;
		a.cover = b.cover;
		a.strand = b.strand;
		a.minAbundance = attr.minAbundance;
		a.timeStamp = attr.timeStamp;
	}

	bool edgeUncovered(const DpAttr &attr, int a, int b) const
	{
		return attr.uncovered.count((long long)a * seCnt + b) != 0;
	}

	// cover evaluation shared by the three places that scan the constraints
	// (.cpp:463-490, 569-596, 732-755).  Returns the new cover; sets belowMin when a compatible
	// constraint has normAbund <= minAbundance.
	double scanCover(const Transcript &t, const std::vector<Constraint> &tc, int tcStart, int bound, double cover,
	                 const DpAttr &attr, bool &belowMin) const
	{
		int size = (int)tc.size();
		for (int i = tcStart; i < size; ++i) {
			if (tc[i].first > bound)
				break;
			if (compat(t, tc[i]) == 1) {
				if (tc[i].normAbund <= attr.minAbundance) {
					belowMin = true;
					return -2;
				}
				if (tc[i].abundance <= 0)
					continue;
				if (attr.forAbundance) {
					if (tc[i].normAbund < cover || cover == 0)
						cover = tc[i].normAbund;
				} else
					++cover;
			}
		}
		return cover;
	}

	Dp solveSub(std::vector<int> &visit, int vcnt, int strand, std::vector<Constraint> &tc, int tcStart, DpAttr &attr);

	// SearchSubTranscript (TranscriptDecider.cpp:374-640)
	void searchSub(int tag, int strand, const int *parents, int pcnt, Dp &pdp, std::vector<int> &visit, int vcnt,
	               const std::vector<int> &extends, int extendCnt, std::vector<Constraint> &tc, int tcStart,
	               DpAttr &attr)
	{
		bool keepSearch = true;
		bool belowMin = false;
		if ((int)visit.size() <= vcnt) visit.resize((size_t)vcnt + 1);
		visit[vcnt] = tag;
		++vcnt;
		Dp visitdp;
		visitdp.cover = -1;

		Transcript sub;
		sub.v = bits_new(seCnt);
		for (int i = 0; i < pcnt; ++i) bit_set(sub.v, parents[i]);
		sub.first = parents[0];
		for (int i = 0; i < vcnt; ++i) bit_set(sub.v, visit[i]);
		sub.last = visit[vcnt - 1];
		sub.partial = true;

		int i;
		for (i = extendCnt - 1; i >= 0; --i) {
			const Constraint &c = tc[extends[i]];
			if (c.last <= tag || (bit_test(c.v, tag) && compat(sub, c) != 0))
				break;
		}
		extendCnt = i + 1;

		sub.partial = false;
		if (!se[tag].next.empty() && (extendCnt == 0 || tag >= tc[extends[extendCnt - 1]].last)) {
			std::vector<int> v2(visit.begin(), visit.begin() + vcnt);
			visitdp = solveSub(v2, vcnt, strand, tc, tcStart, attr);
			keepSearch = false;
		}

		double cover;
		if (visitdp.cover >= 0) {
			cover = visitdp.cover;
			bits_reset(sub.v);
			for (i = 0; i < pcnt; ++i) bit_set(sub.v, parents[i]);
			for (size_t w = 0; w < sub.v.size(); ++w) sub.v[w] |= visitdp.v[w];
			sub.first = parents[0];
			sub.last = visitdp.last;
			sub.partial = false;
			if (!attr.forAbundance && attr.minAbundance > 0) {
				for (i = 0; i < pcnt - 1; ++i)
					if (edgeUncovered(attr, parents[i], parents[i + 1])) belowMin = true;
				for (i = -1; i < vcnt - 1; ++i) {
					if (i == -1 && pcnt >= 1) {
						if (edgeUncovered(attr, parents[pcnt - 1], visit[0])) belowMin = true;
					} else if (edgeUncovered(attr, visit[i], visit[i + 1]))
						belowMin = true;
				}
			}
			bool bm2 = false;
			cover = scanCover(sub, tc, tcStart, parents[pcnt - 1], cover, attr, bm2);
			if (bm2) belowMin = true;
			if (belowMin && pdp.cover == -1) {
				pdp.cover = -2;
				pdp.v = sub.v;
				pdp.first = sub.first;
				pdp.last = sub.last;
				pdp.strand = strand;
			} else if (cover > pdp.cover) {
				pdp.cover = cover;
				pdp.v = sub.v;
				pdp.first = sub.first;
				pdp.last = sub.last;
				pdp.strand = strand;
			}
		} else if (visitdp.cover == -2 && pdp.cover == -1) {
			bits_reset(sub.v);
			for (i = 0; i < pcnt; ++i) bit_set(sub.v, parents[i]);
			for (size_t w = 0; w < sub.v.size(); ++w) sub.v[w] |= visitdp.v[w];
			sub.first = parents[0];
			sub.last = visitdp.last;
			pdp.cover = -2;
			pdp.v = sub.v;
			pdp.first = sub.first;
			pdp.last = sub.last;
			pdp.strand = strand;
		}

		if (se[tag].canEnd && (visitdp.cover < 0 || attr.forAbundance)) {
			bool below = false;
			bits_reset(sub.v);
			for (i = 0; i < pcnt; ++i) bit_set(sub.v, parents[i]);
			for (i = 0; i < vcnt; ++i) bit_set(sub.v, visit[i]);
			sub.first = parents[0];
			sub.last = visit[vcnt - 1];
			sub.partial = false;
			cover = 0;
			if (attr.forAbundance || attr.minAbundance > 0) {
				for (i = 0; i < pcnt - 1; ++i)
					if (edgeUncovered(attr, parents[i], parents[i + 1])) below = true;
				for (i = -1; i < vcnt - 1; ++i) {
					if (i == -1 && pcnt >= 1) {
						if (edgeUncovered(attr, parents[pcnt - 1], visit[0])) below = true;
					} else if (edgeUncovered(attr, visit[i], visit[i + 1]))
						below = true;
				}
				if (attr.forAbundance && below)
					cover = 1e-6;
			}
			bool bm2 = false;
			cover = scanCover(sub, tc, tcStart, visit[vcnt - 1], cover, attr, bm2);
			if (bm2) below = true;
			if (below && pdp.cover == -1) {
				pdp.cover = -2;
				pdp.v = sub.v;
				pdp.first = sub.first;
				pdp.last = sub.last;
				pdp.strand = strand;
			} else if (cover > pdp.cover) {
				pdp.cover = cover;
				pdp.v = sub.v;
				pdp.first = sub.first;
				pdp.last = sub.last;
				pdp.strand = strand;
			}
		}

		if (keepSearch) {
			for (int b : se[tag].next) {
				if ((same_strand(se[tag].rs, strand) && same_strand(se[b].ls, strand)) ||
				    se[b].start == se[tag].end + 1) {
					int backup = strand;
					if (se[b].start > se[tag].end + 1)
						strand = se[tag].rs;
					searchSub(b, strand, parents, pcnt, pdp, visit, vcnt, extends, extendCnt, tc, tcStart, attr);
					strand = backup;
				}
			}
		}
	}

	int hashKey(const std::vector<int> &visit, int vcnt) const
	{
		int key = 0;
		for (int i = 0; i < vcnt; ++i)
			key = (int)(((long long)key * seCnt + visit[i]) % hashMax); // see note: int product in the reference
		if (key < 0) key += hashMax;
		return key;
	}

	// PickTranscriptsByDP (TranscriptDecider.cpp:860-1201)
	void pickByDP(int iterBound, std::vector<Constraint> &tc, DpAttr &attr, std::vector<Transcript> &all);

	// AugmentTranscripts(limit, extend=false) (TranscriptDecider.cpp:1205-1365)
	void augment(std::vector<Transcript> &all, int limit)
	{
		int size = (int)all.size();
		if (size >= limit) return;
		std::vector<int> idx;
		for (int i = 0; i < size; ++i) {
			bits_ones(all[i].v, idx);
			int cnt = (int)idx.size();
			for (int j = 1; j < cnt; ++j)
				if (se[idx[j]].canStart) {
					Transcript nt;
					nt.first = idx[j];
					nt.last = all[i].last;
					nt.v = all[i].v;
					bits_keep(nt.v, nt.first, nt.last);
					all.push_back(nt);
					if ((int)all.size() >= limit) return;
				}
			for (int j = 0; j < cnt - 1; ++j)
				if (se[idx[j]].canEnd) {
					Transcript nt;
					nt.first = all[i].first;
					nt.last = idx[j];
					nt.v = all[i].v;
					bits_keep(nt.v, nt.first, nt.last);
					all.push_back(nt);
					if ((int)all.size() >= limit) return;
				}
		}
		coalesceTx(all);
	}

	void pickTranscripts(std::vector<Transcript> &all, ConstraintSet &cs, std::vector<Transcript> &picked);
	void abundanceEM(ConstraintSet &cs, std::vector<Transcript> &tx);
	int refine(bool aggressive, std::vector<std::map<int, int>> &chainSupport, std::vector<int> &txptSampleSupport,
	           std::vector<Transcript> &tx, ConstraintSet &cs);
	void output(int compIdx, int chr, int sampleId, const Transcript &t, std::vector<OutTranscript> &out);
};

// SolveSubTranscript (TranscriptDecider.cpp:642-857)
Dp Decider::solveSub(std::vector<int> &visit, int vcnt, int strand, std::vector<Constraint> &tc, int tcStart,
                     DpAttr &attr)
{
	int i, size;
	auto hit = [&](const Dp &d) {
		return d.cover != -1 && d.strand == strand &&
		       (d.timeStamp == attr.timeStamp || (d.minAbundance < attr.minAbundance && d.cover == -2));
	};
	if (vcnt == 1) {
		Dp &d = memoSlot(attr.f1[visit[0]], attr);
		if (hit(d)) return d;
	} else if (vcnt == 2 && attr.hasF2) {
		Dp &d = memoSlot(attr.f2[(size_t)visit[0] * seCnt + visit[1]], attr);
		if (hit(d)) return d;
	} else {
		Dp &d = memoSlot(attr.hash[hashKey(visit, vcnt)], attr);
		if (d.cover != -1 && d.cnt == vcnt && d.strand == strand && d.first == visit[0] &&
		    (d.timeStamp == attr.timeStamp || (d.minAbundance < attr.minAbundance && d.cover == -2))) {
			Bits x = bits_new(seCnt);
			for (i = 0; i < vcnt; ++i) bit_set(x, visit[i]);
			for (size_t w = 0; w < x.size(); ++w) x[w] ^= d.v[w];
			bits_keep(x, visit[0], visit[vcnt - 1]);
			if (bits_zero(x)) return d;
		}
	}
	size = (int)tc.size();
	for (i = tcStart; i < size; ++i)
		if (tc[i].first >= visit[0])
			break;
	tcStart = i;

	Dp visitdp;
	visitdp.v = bits_new(seCnt);
	visitdp.cover = -1;
	Transcript sub;
	sub.v = bits_new(seCnt);
	for (i = 0; i < vcnt; ++i) bit_set(sub.v, visit[i]);
	sub.first = visit[0];
	sub.last = visit[vcnt - 1];
	if (se[visit[vcnt - 1]].canEnd) {
		sub.partial = false;
		bool bm = false;
		double cover = scanCover(sub, tc, tcStart, sub.last, 0, attr, bm);
		visitdp.v = sub.v;
		visitdp.cover = cover;
		visitdp.first = sub.first;
		visitdp.last = sub.last;
		visitdp.strand = strand;
	}
	std::vector<int> extends;
	sub.partial = true;
	for (i = tcStart; i < size; ++i) {
		if (tc[i].first > sub.last)
			break;
		if (compat(sub, tc[i]) == 2)
			extends.push_back(i);
	}
	// qsort by .last (TranscriptDecider.cpp:781-796); glibc's qsort is a stable merge sort here
	std::stable_sort(extends.begin(), extends.end(), [&](int a, int b) { return tc[a].last < tc[b].last; });
	int extendCnt = (int)extends.size();
	std::vector<int> nextv;
	{
		int a = visit[vcnt - 1];
		std::vector<int> parents(visit.begin(), visit.begin() + vcnt);
		for (int b : se[a].next) {
			if ((same_strand(se[a].rs, strand) && same_strand(se[b].ls, strand)) || se[b].start == se[a].end + 1) {
				int backup = strand;
				if (se[b].start > se[a].end + 1)
					strand = se[a].rs;
				searchSub(b, strand, parents.data(), vcnt, visitdp, nextv, 0, extends, extendCnt, tc, tcStart, attr);
				strand = backup;
			}
		}
	}
	if (vcnt == 1) {
		Dp &d = memoSlot(attr.f1[visit[0]], attr);
		setDp(d, visitdp, attr);
		return d;
	} else if (vcnt == 2 && attr.hasF2) {
		Dp &d = memoSlot(attr.f2[(size_t)visit[0] * seCnt + visit[1]], attr);
		setDp(d, visitdp, attr);
		return d;
	} else {
		Dp &d = memoSlot(attr.hash[hashKey(visit, vcnt)], attr);
		setDp(d, visitdp, attr);
		d.cnt = vcnt;
		return d;
	}
}

void Decider::pickByDP(int iterBound, std::vector<Constraint> &tc, DpAttr &attr, std::vector<Transcript> &all)
{
	int i, j, k;
	std::vector<Transcript> transcripts;
	int tcCnt = (int)tc.size();
	attr.timeStamp = 1;
	++attr.epoch; // == ResetDpContent on f1, f2, hash (.cpp:882-888)
	double maxAbundance = -1;

	// uncovered edges (.cpp:890-947)
	attr.uncovered.clear();
	k = 0;
	for (i = 0; i < seCnt; ++i) {
		for (; k < tcCnt; ++k)
			if (tc[k].last >= i)
				break;
		if (k >= tcCnt || tc[k].first > i) {
			for (int n : se[i].next) attr.uncovered.insert((long long)i * seCnt + n);
			continue;
		}
		for (int n : se[i].next) {
			bool covered = false;
			for (int l = k; l < tcCnt; ++l) {
				if (tc[l].first > i)
					break;
				if (bit_test(tc[l].v, i) && bit_test(tc[l].v, n)) {
					if (n == i + 1) {
						covered = true;
						break;
					}
					Bits b = tc[l].v;
					bits_keep(b, i + 1, n - 1);
					if (bits_zero(b)) {
						covered = true;
						break;
					}
				}
			}
			if (!covered)
				attr.uncovered.insert((long long)i * seCnt + n);
		}
	}

	attr.forAbundance = true;
	attr.minAbundance = 0;
	std::vector<int> visit(1);
	for (i = 0; i < seCnt; ++i)
		if (se[i].canStart) {
			visit.assign(1, i);
			Dp tmp = solveSub(visit, 1, 0, tc, 0, attr);
			if (tmp.cover > maxAbundance)
				maxAbundance = tmp.cover;
		}

	attr.forAbundance = false;
	std::vector<int> coveredTc;
	Dp maxCoverDp, bestDp;
	maxCoverDp.v = bits_new(seCnt);
	bestDp.v = bits_new(seCnt);
	std::map<double, Dp> cached;
	int iterCnt = 0;
	Transcript sub;
	sub.v = bits_new(seCnt);
	while (1) {
		double bestScore;
		attr.minAbundance = 0;
		bestDp.cover = -1;
		bestScore = -1;
		while (1) {
			auto cit = cached.find(attr.minAbundance);
			if (cit != cached.end()) {
				Dp tmp = cit->second;
				setDp(maxCoverDp, tmp, attr);
			} else {
				maxCoverDp.cover = -1;
				++attr.timeStamp;
				for (i = 0; i < seCnt; ++i) {
					if (!se[i].canStart)
						continue;
					visit.assign(1, i);
					Dp tmp = solveSub(visit, 1, 0, tc, 0, attr);
					if (tmp.cover > maxCoverDp.cover && tmp.cover > 0)
						setDp(maxCoverDp, tmp, attr);
				}
				if (maxCoverDp.cover == -1)
					break;
				Dp ccr;
				setDp(ccr, maxCoverDp, attr);
				cached[attr.minAbundance] = ccr;
			}
			double min = -1;
			sub.v = maxCoverDp.v;
			sub.first = maxCoverDp.first;
			sub.last = maxCoverDp.last;
			// NOTE: sub.partial keeps whatever the last user of the shared buffer left (the
			// reference's attr.bufferTxpt); every path leaves it false here.
			sub.partial = false;
			for (i = 0; i < tcCnt; ++i)
				if (compat(sub, tc[i]) == 1)
					if (tc[i].normAbund < min || min == -1)
						min = tc[i].normAbund;
			if (attr.minAbundance == 0) {
				std::vector<int> idx;
				bits_ones(maxCoverDp.v, idx);
				for (i = 0; i + 1 < (int)idx.size(); ++i)
					if (edgeUncovered(attr, idx[i], idx[i + 1])) {
						min = 1e-6;
						break;
					}
			}
			double sc = score(maxCoverDp.cover, 1.0, min, maxAbundance, 0);
			if (bestScore == -1 || sc > bestScore) {
				bestScore = sc;
				setDp(bestDp, maxCoverDp, attr);
			} else if (sc < bestScore) {
				if (score(maxCoverDp.cover, 1.0, maxAbundance, maxAbundance, 0) < bestScore)
					break;
			}
			attr.minAbundance = min;
		}
		if (bestDp.cover == -1)
			break;
		coveredTc.clear();
		double update = -1;
		sub.v = bestDp.v;
		sub.first = bestDp.first;
		sub.last = bestDp.last;
		sub.partial = false;
		for (i = 0; i < tcCnt; ++i)
			if (compat(sub, tc[i]) == 1) {
				if (tc[i].abundance > 0 && (tc[i].abundance < update || update == -1))
					update = tc[i].abundance;
				coveredTc.push_back(i);
			}
		update *= (1 + iterCnt / 50); // Q7: integer division

		Transcript nt;
		nt.v = bestDp.v;
		nt.first = bestDp.first;
		nt.last = bestDp.last;
		nt.partial = false;
		nt.abundance = 0;
		for (i = 0; i < (int)coveredTc.size(); ++i) {
			j = coveredTc[i];
			if (tc[j].abundance > 0) {
				double tmp = (tc[j].abundance > update ? update : tc[j].abundance);
				tc[j].abundance -= tmp;
				double factor = 1;
				nt.abundance += (tc[j].support * update / tc[j].normAbund * factor);
				if (tc[j].abundance <= 0) {
					std::vector<double> removeKey;
					for (auto &kv : cached) {
						Transcript ct;
						ct.v = kv.second.v;
						ct.first = kv.second.first;
						ct.last = kv.second.last;
						ct.partial = false;
						if (compat(ct, tc[j]) == 1)
							removeKey.push_back(kv.first);
					}
					for (double key : removeKey) cached.erase(key);
				}
			}
			if (tc[j].abundance < 0)
				tc[j].abundance = 0;
		}
		transcripts.push_back(nt);
		// the capacity-triggered partial coalescing (.cpp:1155-1160) only reorders equal
		// structures before the final CoalesceSameTranscripts; abundance sums are not used later
		++iterCnt;
		if (iterCnt >= iterBound)
			break;
	}
	coalesceTx(transcripts);
	for (auto &t : transcripts) {
		t.corr = 0; // correlation.Query()==0 for every pair -> cor=0; single-subexon: cor=2 -> 0 (.cpp:1169-1185)
		all.push_back(t);
	}
}

// PickTranscripts (TranscriptDecider.cpp:1368-2119)
void Decider::pickTranscripts(std::vector<Transcript> &all, ConstraintSet &cs, std::vector<Transcript> &picked)
{
	int i, j, k;
	std::vector<MatePair> &tc = cs.p;
	int atcnt = (int)all.size();
	int tcCnt = (int)tc.size();
	if (tcCnt == 0 || atcnt == 0)
		return;
	double inf = -1;
	std::vector<int> transcriptLength((size_t)atcnt);
	std::vector<double> transcriptAbundance((size_t)atcnt), avgTranscriptAbundance((size_t)atcnt, 0.0);
	std::vector<double> coveredPortion((size_t)atcnt);
	std::vector<Bits> btable((size_t)atcnt, bits_new(tcCnt));
	for (j = 0; j < tcCnt; ++j) {
		int a = tc[j].i, b = tc[j].j;
		if (cs.c[a].support > inf) inf = cs.c[a].support;
		if (cs.c[b].support > inf) inf = cs.c[b].support;
		if (tc[j].normAbund > inf) inf = tc[j].normAbund;
		tc[j].abundance = tc[j].normAbund;
	}
	++inf;
	bool btableSet = false;
	for (i = 0; i < atcnt; ++i)
		for (j = 0; j < tcCnt; ++j)
			if (compat(all[i], cs.c[tc[j].i]) == 1 && compat(all[i], cs.c[tc[j].j]) == 1) {
				bit_set(btable[i], j);
				btableSet = true;
			}
	if (!btableSet)
		return;

	double maxAbundance = -1;
	{
		std::vector<std::pair<int, int>> chain((size_t)seCnt);
		std::vector<char> covered((size_t)seCnt);
		std::vector<char> usedConstraints(cs.c.size());
		std::vector<Bits> togetherChain;
		auto isTogether = [&](int q) {
			return (se[q].lt == 2 && se[q].rt == 1) || (se[q].lt == 0 && se[q].rt == 1) ||
			       (se[q].lt == 2 && se[q].rt == 0);
		};
		for (i = 0; i < seCnt;) {
			for (j = i + 1; j < seCnt; ++j)
				if (se[j].start > se[j - 1].end + 1)
					break;
			int cnt = 0;
			for (k = i; k < j; ++k)
				if (isTogether(k)) ++cnt;
			if (cnt <= 1) {
				i = j;
				continue;
			}
			Bits tmp = bits_new(seCnt);
			for (k = i; k < j; ++k)
				if (isTogether(k)) bit_set(tmp, k);
			togetherChain.push_back(tmp);
			i = j;
		}
		std::vector<int> idx;
		for (i = 0; i < atcnt; ++i) {
			double value = inf;
			int tag = -1;
			all[i].abundance = 0;
			all[i].csupport.assign((size_t)tcCnt, 0.0);
			bits_ones(all[i].v, idx);
			int seIdxCnt = (int)idx.size();
			transcriptLength[i] = 0;
			for (j = 0; j < seIdxCnt; ++j)
				transcriptLength[i] += se[idx[j]].end - se[idx[j]].start + 1;
			for (j = 0; j < seIdxCnt - 1; ++j) {
				chain[j] = {idx[j], idx[j + 1]};
				covered[j] = 0;
			}
			std::fill(usedConstraints.begin(), usedConstraints.end(), 0);
			int compatibleCnt = 0;
			for (j = 0; j < tcCnt; ++j) {
				all[i].csupport[j] = 0;
				if (bit_test(btable[i], j) && tc[j].abundance > 0) {
					++compatibleCnt;
					double adjustAbundance = tc[j].abundance;
					const Constraint &ci = cs.c[tc[j].i], &cj = cs.c[tc[j].j];
					if (seIdxCnt > 1) {
						bool iEdge = ci.first + ci.last == 2 * all[i].first || ci.first + ci.last == 2 * all[i].last;
						bool jEdge = cj.first + cj.last == 2 * all[i].first || cj.first + cj.last == 2 * all[i].last;
						if (tc[j].i == tc[j].j && iEdge)
							adjustAbundance = inf;
						else if (tc[j].i != tc[j].j && iEdge)
							adjustAbundance = cj.normAbund;
						else if (tc[j].i != tc[j].j && jEdge)
							adjustAbundance = ci.normAbund;
					}
					if (adjustAbundance < value) {
						value = adjustAbundance;
						tag = j;
					}
					avgTranscriptAbundance[i] += tc[j].abundance;
					if (!usedConstraints[tc[j].i]) {
						for (k = 0; k < seIdxCnt - 1; ++k)
							if (bit_test(ci.v, chain[k].first) && bit_test(ci.v, chain[k].second))
								covered[k] = 1;
						usedConstraints[tc[j].i] = 1;
					}
					if (!usedConstraints[tc[j].j]) {
						for (k = 0; k < seIdxCnt - 1; ++k)
							if (bit_test(cj.v, chain[k].first) && bit_test(cj.v, chain[k].second))
								covered[k] = 1;
						usedConstraints[tc[j].j] = 1;
					}
				}
			}
			for (auto &tg : togetherChain) {
				Bits b = tg;
				for (size_t w = 0; w < b.size(); ++w) b[w] &= all[i].v[w];
				if (!bits_zero(b) && !bits_equal(b, tg))
					value /= 2;
			}
			int cnt = 0;
			for (j = 0; j < seIdxCnt - 1; ++j) {
				if (!covered[j])
					value = 0;
				else
					++cnt;
			}
			if (seIdxCnt > 1)
				coveredPortion[i] = (double)cnt / (double)(seIdxCnt - 1);
			else
				coveredPortion[i] = 1;
			if (coveredPortion[i] == 0)
				coveredPortion[i] = (double)0.5 / (seIdxCnt);
			if (tag == -1)
				value = 0;
			if (value > maxAbundance)
				maxAbundance = value;
			transcriptAbundance[i] = value;
			if (tag != -1)
				avgTranscriptAbundance[i] /= compatibleCnt;
		}
		if (maxAbundance == 0) {
			for (i = 0; i < atcnt; ++i)
				transcriptAbundance[i] = coveredPortion[i];
			maxAbundance = 1;
		}
	}

	// quantitative set cover (.cpp:1827-2068)
	std::vector<double> coverCnt((size_t)atcnt, -1.0);
	std::vector<int> list((size_t)atcnt);
	while (1) {
		double max = -1;
		int maxtag = -1;
		double maxcnt = -1;
		for (i = 0; i < atcnt; ++i) {
			double value, cnt = 0;
			if (coverCnt[i] == -1) {
				for (j = 0; j < tcCnt; ++j)
					if (tc[j].abundance > 0 && bit_test(btable[i], j))
						cnt += tc[j].effectiveCount;
				coverCnt[i] = cnt;
			} else
				cnt = coverCnt[i];
			value = transcriptAbundance[i];
			if (cnt < 1)
				continue;
			cnt *= coveredPortion[i];
			double weight = 1;
			if (all[i].FPKM > 0 && prm.sampleCnt > 1)
				weight = (1 + all[i].FPKM / prm.sampleCnt);
			double sc = score(cnt, weight, value, maxAbundance, all[i].corr);
			if (cnt > maxcnt)
				maxcnt = cnt;
			if (sc > max) {
				max = sc;
				maxtag = i;
			} else if (sc == max) {
				if (avgTranscriptAbundance[maxtag] < avgTranscriptAbundance[i]) {
					max = sc;
					maxtag = i;
				}
			}
		}
		if (maxcnt == 0 || maxtag == -1)
			break;
		double update = inf;
		int updateTag = 0;
		for (j = 0; j < tcCnt; ++j)
			if (bit_test(btable[maxtag], j) && tc[j].abundance > 0 && tc[j].abundance <= update) {
				update = tc[j].abundance;
				updateTag = j;
			}
		int supportCnt = 0;
		for (j = 0; j < tcCnt; ++j) {
			if (bit_test(btable[maxtag], j)) {
				if (tc[j].abundance > 0) {
					tc[j].abundance -= 1 * update;
					double factor = tc[j].effectiveCount;
					double tmp = (tc[j].support * update / tc[j].normAbund * factor);
					all[maxtag].csupport[j] += tmp;
					all[maxtag].abundance += tmp;
					if (tc[j].abundance <= 0)
						for (int l = 0; l < atcnt; ++l)
							if (bit_test(btable[l], j))
								coverCnt[l] -= tc[j].effectiveCount;
					++supportCnt;
				} else if (all[maxtag].csupport[j] == 0) {
					double sum = 0, takeOut = 0;
					double factor = tc[j].effectiveCount;
					int listCnt = 0;
					for (i = 0; i < atcnt; ++i) {
						if (i == maxtag)
							continue;
						if (all[i].abundance > 0 && bit_test(btable[i], j)) {
							sum += all[i].csupport[j];
							double tmp = (all[i].csupport[j] + all[maxtag].csupport[j]) * transcriptAbundance[maxtag] /
							                 (transcriptAbundance[maxtag] + transcriptAbundance[i]) -
							             all[maxtag].csupport[j];
							if (tmp > 0) {
								list[listCnt++] = i;
								takeOut += tmp;
							}
						}
					}
					double ratio = 1, takeOutFactor = 0.5;
					if (update < tc[j].normAbund) {
						if (takeOut > (tc[j].support * update / tc[j].normAbund * factor) * takeOutFactor)
							ratio = (tc[j].support * update / tc[j].normAbund * factor) * takeOutFactor / takeOut;
					} else {
						if (takeOut > (tc[j].support * factor) * takeOutFactor)
							ratio = (tc[j].support * factor) * takeOutFactor / takeOut;
					}
					for (i = 0; i < listCnt; ++i) {
						double tmp = ((all[list[i]].csupport[j] + all[maxtag].csupport[j]) * transcriptAbundance[maxtag] /
						                  (transcriptAbundance[maxtag] + transcriptAbundance[list[i]]) -
						              all[maxtag].csupport[j]) *
						             ratio;
						all[list[i]].csupport[j] -= tmp;
						all[list[i]].abundance -= tmp;
					}
					double tmp = takeOut * ratio;
					all[maxtag].csupport[j] += tmp;
					all[maxtag].abundance += tmp;
				}
			}
			if (tc[j].abundance < 0)
				tc[j].abundance = 0;
		}
		tc[updateTag].abundance = 0;
		if (supportCnt == 0)
			break;
	}
	for (i = 0; i < atcnt; ++i)
		if (all[i].abundance > 0) {
			Transcript nt = all[i];
			nt.csupport.clear();
			nt.corr = transcriptAbundance[i] == 0 ? -1 : 0;
			nt.id = i;
			picked.push_back(nt);
		}
	for (i = 0; i < atcnt; ++i)
		all[i].csupport.clear();
}

// AbundanceEstimation (TranscriptDecider.cpp:2121-2251)
void Decider::abundanceEM(ConstraintSet &cs, std::vector<Transcript> &tx)
{
	int tcnt = (int)tx.size();
	if (tcnt <= 0) return;
	std::vector<MatePair> &tc = cs.p;
	int tcCnt = (int)tc.size();
	std::vector<Bits> btable((size_t)tcnt, bits_new(tcCnt));
	std::vector<int> length((size_t)tcnt), compatibleList((size_t)tcnt);
	std::vector<double> rho((size_t)tcnt);
	int i, j, iterCnt = 0;
	for (i = 0; i < tcnt; ++i) {
		tx[i].csupport.assign((size_t)tcCnt, 0.0);
		for (j = 0; j < tcCnt; ++j)
			if (compat(tx[i], cs.c[tc[j].i]) == 1 && compat(tx[i], cs.c[tc[j].j]) == 1)
				bit_set(btable[i], j);
		int len = txLen(tx[i].v);
		length[i] = len - prm.fragLen + 2 * prm.fragStdev;
		if (length[i] < 1) length[i] = 1;
		rho[i] = tx[i].abundance / length[i];
		if (tx[i].corr == -1 && rho[i] > 0.1 / (double)prm.readLen)
			rho[i] = 0.1 / (double)prm.readLen;
	}
	while (1) {
		for (i = 0; i < tcnt; ++i)
			std::fill(tx[i].csupport.begin(), tx[i].csupport.end(), 0.0);
		for (j = 0; j < tcCnt; ++j) {
			int clCnt = 0;
			double sum = 0;
			for (i = 0; i < tcnt; ++i)
				if (bit_test(btable[i], j)) {
					compatibleList[clCnt++] = i;
					sum += rho[i];
				}
			for (i = 0; i < clCnt; ++i) {
				double factor = tc[j].effectiveCount;
				tx[compatibleList[i]].csupport[j] = (rho[compatibleList[i]] / sum) * tc[j].support * factor;
			}
		}
		double diff = 0;
		for (i = 0; i < tcnt; ++i) {
			double newAbund = 0;
			for (j = 0; j < tcCnt; ++j)
				newAbund += tx[i].csupport[j];
			double old = rho[i];
			rho[i] = newAbund / length[i];
			if (tx[i].corr == -1 && rho[i] > 0.1 / (double)prm.readLen)
				rho[i] = 0.1 / (double)prm.readLen;
			double tmp = (old - rho[i]);
			diff += tmp < 0 ? -tmp : tmp;
		}
		if (diff < 1e-3)
			break;
		++iterCnt;
		if (iterCnt >= 1000)
			break;
	}
	for (i = 0; i < tcnt; ++i) {
		tx[i].abundance = 0;
		for (j = 0; j < tcCnt; ++j)
			tx[i].abundance += tx[i].csupport[j];
		tx[i].csupport.clear();
	}
}

static int removeNegative(std::vector<Transcript> &tx)
{
	size_t j = 0;
	for (size_t i = 0; i < tx.size(); ++i) {
		if (tx[i].abundance < 0)
			continue;
		if (i != j) tx[j] = tx[i];
		++j;
	}
	tx.resize(j);
	return (int)j;
}

// RefineTranscripts (TranscriptDecider.cpp:2253-2837)
int Decider::refine(bool aggressive, std::vector<std::map<int, int>> &chainSupport, std::vector<int> &txptSampleSupport,
                    std::vector<Transcript> &tx, ConstraintSet &cs)
{
	int i, j, k;
	int tcnt = (int)tx.size();
	if (tcnt == 0)
		return 0;
	std::vector<MatePair> &tc = cs.p;
	std::vector<Constraint> &scc = cs.c;
	int tcCnt = (int)tc.size();
	int ng = usedGeneId - baseGeneId;
	std::vector<double> geneMaxFPKM((size_t)ng, 0.0), nonOverlapMaxFPKM((size_t)ng, 0.0), geneMaxCov((size_t)ng, 0.0);
	std::vector<int> geneMaxFPKMTag((size_t)ng, 0), txptGid((size_t)tcnt);
	int readLen = prm.readLen;

	for (i = 0; i < tcnt; ++i) {
		int gid = geneIdOf(tx[i]);
		int len = lenFromFPKM(tx[i].abundance, tx[i].FPKM);
		if (tx[i].FPKM > geneMaxFPKM[gid - baseGeneId]) {
			geneMaxFPKM[gid - baseGeneId] = tx[i].FPKM;
			geneMaxFPKMTag[gid - baseGeneId] = i;
		}
		if (tx[i].abundance * readLen / len > geneMaxCov[gid - baseGeneId])
			geneMaxCov[gid - baseGeneId] = (tx[i].abundance * readLen) / len;
		txptGid[i] = gid;
	}
	for (i = 0; i < tcnt; ++i) {
		int tag = txptGid[i] - baseGeneId;
		if ((tx[i].last < tx[geneMaxFPKMTag[tag]].first || tx[i].first > tx[geneMaxFPKMTag[tag]].last) &&
		    tx[i].FPKM > nonOverlapMaxFPKM[tag])
			nonOverlapMaxFPKM[tag] = tx[i].FPKM;
	}
	if (!aggressive) {
		int cnt, tag = 0;
		std::vector<int> uniqCount((size_t)tcnt, 0);
		for (j = 0; j < tcCnt; ++j) {
			cnt = 0;
			if (tc[j].uniq <= 5)
				continue;
			for (i = 0; i < tcnt; ++i) {
				if (compat(tx[i], scc[tc[j].i]) && compat(tx[i], scc[tc[j].j])) {
					tag = i;
					++cnt;
				}
				if (cnt >= 2)
					break;
			}
			if (cnt == 1)
				++uniqCount[tag];
		}
		for (i = 0; i < tcnt; ++i)
			if (uniqCount[i] >= 2) {
				tx[i].abundance *= 4;
				tx[i].FPKM *= 4;
			}
	}
	int sccCnt = (int)scc.size();
	double filterFactor = 1.0;
	for (i = 0; i < tcnt; ++i)
		if (tx[i].FPKM < filterFactor * prm.fpkmFraction * geneMaxFPKM[txptGid[i] - baseGeneId])
			tx[i].abundance = -tx[i].abundance;
	{ // Q6: "if ( nonOverlapMaxFPKM != 0 )" tests a pointer -> always taken
		std::vector<int> rescueList;
		for (i = 0; i < tcnt; ++i) {
			if (tx[i].abundance >= 0)
				continue;
			for (j = 0; j < tcnt; ++j) {
				if (tx[j].abundance < 0 || txptGid[i] != txptGid[j])
					continue;
				if (tx[i].first <= tx[j].last && tx[i].last >= tx[j].first)
					break;
			}
			if (j >= tcnt && tx[i].FPKM >= prm.fpkmFraction * nonOverlapMaxFPKM[txptGid[i] - baseGeneId])
				rescueList.push_back(i);
		}
		for (int r : rescueList)
			tx[r].abundance *= -1;
	}
	// coverage filter (-d)
	Bits buffer = bits_new(seCnt);
	for (i = 0; i < tcnt; ++i) {
		if (tx[i].abundance >= 0) {
			int len = lenFromFPKM(tx[i].abundance, tx[i].FPKM);
			double cov = (tx[i].abundance * readLen) / len;
			if ((tcnt > 1 || len <= 1000 || bits_count(tx[i].v) <= 3) && cov < prm.minDepth) {
				if (geneMaxCov[txptGid[i] - baseGeneId] == cov)
					continue;
				if (geneMaxCov[txptGid[i] - baseGeneId] <= 5) {
					bits_reset(buffer);
					for (j = 0; j < sccCnt; ++j) {
						if (!compat(tx[i], scc[j]))
							continue;
						for (size_t w = 0; w < buffer.size(); ++w) buffer[w] |= scc[j].v[w];
					}
					if (bits_equal(buffer, tx[i].v))
						tx[i].abundance = -tx[i].abundance;
				} else
					tx[i].abundance = -tx[i].abundance;
			}
		}
	}
	// too short
	for (i = 0; i < tcnt; ++i) {
		if (tx[i].abundance <= 0)
			continue;
		int len = lenFromFPKM(tx[i].abundance, tx[i].FPKM);
		if (len < 200)
			tx[i].abundance = -tx[i].abundance;
	}
	// rescue through cross-sample subexon-chain support (.cpp:2508-2706)
	Constraint tmpC;
	tmpC.v = bits_new(seCnt);
	std::vector<std::pair<int, int>> missingChain;
	std::vector<int> recoverCandidate;
	std::vector<char> used((size_t)tcnt, 0);
	for (i = 0; i < seCnt && prm.sampleCnt > 1; ++i) {
		for (auto it = chainSupport[i].begin(); it != chainSupport[i].end(); ++it) {
			if (prm.sampleCnt >= 0 && (it->second < 3 || it->second < (int)(0.5 * prm.sampleCnt)) &&
			    it->second <= prm.sampleCnt / 2)
				continue;
			bool recover = true;
			bits_reset(tmpC.v);
			bit_set(tmpC.v, i);
			bit_set(tmpC.v, it->first);
			tmpC.first = i;
			tmpC.last = it->first;
			for (j = 0; j < tcnt; ++j) { // Q5: the inner loop reuses j
				if (tx[j].abundance < 0)
					continue;
				if (compat(tx[j], tmpC)) {
					recover = false;
					break;
				}
				if (recover) {
					for (j = 0; j < tcnt; ++j) {
						if (tx[j].abundance > 0)
							continue;
						if (compat(tx[j], tmpC)) {
							missingChain.push_back({i, it->first});
							if (!used[j]) {
								recoverCandidate.push_back(j);
								used[j] = 1;
							}
						}
					}
				}
			}
		}
	}
	int size = (int)recoverCandidate.size();
	std::fill(used.begin(), used.end(), 0);
	std::vector<int> geneRecoverCnt((size_t)ng, 0);
	int round = 1;
	for (i = 0; i < size; ++i) {
		int maxTag = -1;
		for (j = 0; j < size; ++j) {
			if (!used[recoverCandidate[j]]) {
				if (maxTag == -1) {
					maxTag = recoverCandidate[j];
					continue;
				}
				if (txptSampleSupport[tx[recoverCandidate[j]].id] > txptSampleSupport[tx[maxTag].id])
					maxTag = recoverCandidate[j];
				else if (txptSampleSupport[tx[recoverCandidate[j]].id] == txptSampleSupport[tx[maxTag].id]) {
					if (tx[recoverCandidate[j]].FPKM > tx[maxTag].FPKM)
						maxTag = recoverCandidate[j];
				}
			}
		}
		if (maxTag == -1 || txptSampleSupport[tx[maxTag].id] <= 2 ||
		    txptSampleSupport[tx[maxTag].id] < 0.5 * prm.sampleCnt)
			break;
		used[maxTag] = 1;
		if (geneRecoverCnt[txptGid[maxTag] - baseGeneId] >= round)
			continue;
		++geneRecoverCnt[txptGid[maxTag] - baseGeneId];
		k = (int)missingChain.size();
		int cnt = 0;
		for (j = 0; j < k; ++j) {
			if (missingChain[j].first == -1)
				continue;
			bits_reset(tmpC.v);
			bit_set(tmpC.v, missingChain[j].first);
			bit_set(tmpC.v, missingChain[j].second);
			tmpC.first = missingChain[j].first;
			tmpC.last = missingChain[j].second;
			if (compat(tx[maxTag], tmpC)) {
				missingChain[j].first = -1;
				++cnt;
			}
		}
		int len = lenFromFPKM(tx[maxTag].abundance, tx[maxTag].FPKM);
		double cov = (tx[maxTag].abundance * readLen) / len;
		if (cnt >= 1 && cov > 1.0)
			tx[maxTag].abundance *= -1;
	}
	tcnt = removeNegative(tx);

	// low uniqueness
	for (i = 0; i < tcnt; ++i) {
		int support = 0, uniqSupport = 0;
		for (j = 0; j < tcCnt; ++j) {
			if (!compat(tx[i], scc[tc[j].i]) || !compat(tx[i], scc[tc[j].j]))
				continue;
			support += tc[j].support;
			uniqSupport += tc[j].uniq;
		}
		if ((double)uniqSupport < 0.03 * support)
			tx[i].abundance = -1;
	}
	tcnt = removeNegative(tx);

	// shadow 2-exon transcripts (.cpp:2749-2834)
	int minusCnt = 0, plusCnt = 0, mainStrand;
	for (i = 0; i < seCnt; ++i) {
		if (se[i].rs == 1) ++plusCnt;
		else if (se[i].rs == -1) ++minusCnt;
	}
	mainStrand = plusCnt > minusCnt ? 1 : -1;
	std::vector<int> idx;
	for (i = 0; i < tcnt; ++i) {
		bits_ones(tx[i].v, idx);
		int sz = (int)idx.size();
		int intronCnt = 0, anchorIdx = 0;
		for (j = 0; j < sz - 1; ++j)
			if (se[idx[j]].end + 1 < se[idx[j + 1]].start) {
				++intronCnt;
				anchorIdx = j;
			}
		if (intronCnt != 1)
			continue;
		int anchorExonLength[2] = {0, 0};
		int tag = 0;
		for (j = 0; j < sz; ++j) {
			anchorExonLength[tag] += se[idx[j]].end - se[idx[j]].start + 1;
			if (tag == 0 && se[idx[j]].end + 1 < se[idx[j + 1]].start)
				++tag;
		}
		int flag = 0;
		if (se[idx[anchorIdx]].rs == mainStrand) {
			j = idx[anchorIdx];
			if (se[j].end - se[j].start + 1 <= 20 ||
			    (se[j + 1].start == se[j].end + 1 && se[j + 1].end - se[j + 1].start + 1 <= 20 &&
			     se[j + 1].rs == mainStrand))
				++flag;
			j = idx[anchorIdx + 1];
			if (se[j].end - se[j].start + 1 <= 20 ||
			    (se[j].start == se[j - 1].end + 1 && se[j - 1].end - se[j - 1].start + 1 <= 20 &&
			     se[j - 1].ls == mainStrand))
				++flag;
		}
		if (anchorExonLength[0] <= 25 || anchorExonLength[1] <= 25)
			flag = 2;
		int support = 0, uniqSupport = 0;
		for (j = 0; j < tcCnt; ++j) {
			if (!compat(tx[i], scc[tc[j].i]) || !compat(tx[i], scc[tc[j].j]))
				continue;
			if ((bit_test(scc[tc[j].i].v, idx[anchorIdx]) && bit_test(scc[tc[j].i].v, idx[anchorIdx + 1])) ||
			    (bit_test(scc[tc[j].j].v, idx[anchorIdx]) && bit_test(scc[tc[j].j].v, idx[anchorIdx + 1]))) {
				support += tc[j].support;
				uniqSupport += tc[j].uniq;
			}
		}
		if ((double)uniqSupport < 0.3 * support || support < prm.minDepth)
			flag = 2;
		if (flag == 2)
			tx[i].abundance = -1;
	}
	tcnt = removeNegative(tx);
	return (int)tx.size();
}

// OutputTranscript (TranscriptDecider.cpp:3-111)
void Decider::output(int compIdx, int chr, int sampleId, const Transcript &t, std::vector<OutTranscript> &out)
{
	std::vector<int> ind;
	bits_ones(t.v, ind);
	int size = (int)ind.size();
	char strand = '.';
	for (int i = 0; i < size - 1; ++i)
		if (se[ind[i]].end + 1 < se[ind[i + 1]].start) {
			if (se[ind[i]].rs == 1) strand = '+';
			else if (se[ind[i]].rs == -1) strand = '-';
			break;
		}
	OutTranscript o;
	o.comp = compIdx;
	o.sample = sampleId;
	o.chr = chr;
	o.geneId = geneIdOf(ind);
	o.strand = strand;
	o.exons.push_back({se[ind[0]].start + 1, se[ind[0]].end + 1});
	for (int i = 1; i < size; ++i) {
		if (se[ind[i]].start + 1 == o.exons.back().second + 1)
			o.exons.back().second = se[ind[i]].end + 1;
		else
			o.exons.push_back({se[ind[i]].start + 1, se[ind[i]].end + 1});
	}
	int len = 0;
	for (auto &e : o.exons) len += e.second - e.first + 1;
	o.fpkm = t.FPKM;
	o.cov = t.abundance * prm.readLen / len;
	out.push_back(o);
}

} // namespace

// TranscriptDecider::Solve (TranscriptDecider.cpp:2874-3460)
void solve(int compIdx, Component &comp, std::vector<ConstraintSet> &cs, const Params &prm,
           std::vector<OutTranscript> &out, SolveStats *stats)
{
	std::vector<Subexon> &se = comp.se;
	int seCnt = (int)se.size();
	int sampleCnt = prm.sampleCnt;
	int i, j, k;
	Decider D(prm, se);
	double thr = prm.classifierThreshold;

	// start / end flags (.cpp:2884-2973)
	for (i = 0; i < seCnt; ++i) {
		se[i].canStart = se[i].canEnd = false;
		if (se[i].prev.empty())
			se[i].canStart = true;
		else if (se[i].lcls < thr && se[i].lcls != -1 && se[i].ls != 0) {
			bool flag = true;
			for (j = i - 1; j >= 0; --j) {
				if (se[j].end + 1 != se[j + 1].start)
					break;
				if (se[i].canStart == true) { // Q12: tests [i] (just cleared), never true
					flag = false;
					break;
				}
			}
			se[i].canStart = flag;
		}
		if (se[i].next.empty())
			se[i].canEnd = true;
		else if (se[i].rcls < thr && se[i].rcls != -1 && se[i].rs != 0)
			se[i].canEnd = true;
		if (se[i].canEnd) {
			for (j = i - 1; j >= 0; --j) {
				if (se[j].end + 1 != se[j + 1].start)
					break;
				if (se[j].canEnd) {
					se[j].canEnd = false;
					break;
				}
			}
		}
	}
	for (i = 0; i < seCnt;) {
		for (j = i + 1; j < seCnt; ++j)
			if (se[j].start > se[j - 1].end + 1)
				break;
		if (!se[i].canStart) {
			int leftStrandCnt[2] = {0, 0};
			for (k = i; k < j; ++k) {
				if (!same_strand(se[k].rs, se[i].ls))
					break;
				if (se[k].ls != 0)
					++leftStrandCnt[(se[k].ls + 1) / 2];
			}
			if (k < j && leftStrandCnt[(se[k].rs + 1) / 2] == 0)
				se[i].canStart = true;
		}
		if (!se[j - 1].canEnd) {
			int rightStrandCnt[2] = {0, 0};
			for (k = j - 1; k >= i; --k) {
				if (!same_strand(se[k].ls, se[j - 1].rs))
					break;
				if (se[k].rs != 0)
					++rightStrandCnt[(se[k].rs + 1) / 2];
			}
			if (k >= i && rightStrandCnt[(se[k].ls + 1) / 2] == 0)
				se[j - 1].canEnd = true;
		}
		i = j;
	}

	// gene ids (.cpp:2979-3003)
	D.baseGeneId = comp.baseGeneId;
	D.usedGeneId = comp.usedGeneId;
	D.defaultGeneId[0] = D.defaultGeneId[1] = -1;
	for (i = 0; i < seCnt; ++i) {
		if (se[i].geneId < 0)
			continue;
		if ((se[i].rs == -1 || se[i].ls == -1) && (D.defaultGeneId[0] == -1 || se[i].geneId < D.defaultGeneId[0]))
			D.defaultGeneId[0] = se[i].geneId;
		if ((se[i].rs == 1 || se[i].ls == 1) && (D.defaultGeneId[1] == -1 || se[i].geneId < D.defaultGeneId[1]))
			D.defaultGeneId[1] = se[i].geneId;
	}
	if (D.defaultGeneId[0] == -1) D.defaultGeneId[0] = D.baseGeneId;
	if (D.defaultGeneId[1] == -1) D.defaultGeneId[1] = D.usedGeneId - 1;

	// cross-sample subexon-chain support (.cpp:3005-3064)
	std::vector<std::map<int, int>> chainSupport((size_t)seCnt);
	std::vector<int> idx;
	for (i = 0; i < sampleCnt; ++i) {
		std::set<std::pair<int, int>> chain;
		for (auto &c : cs[i].c) {
			if (c.uniq < 0.95 * c.support || c.support < 3)
				continue;
			bits_ones(c.v, idx);
			for (k = 0; k + 1 < (int)idx.size(); ++k)
				if (se[idx[k]].end + 1 < se[idx[k + 1]].start)
					chain.insert({idx[k], idx[k + 1]});
		}
		// NB the reference's dedup leaves one element even when `chain` is empty (resize(k+1) with
		// size 0 -> a default {0,0}... see test_oracle notes); handled below.
		for (auto &p : chain)
			++chainSupport[p.first][p.second];
		if (chain.empty()) {
			// chain.resize(0 + 1) on an empty vector value-initialises one pair {0,0} (.cpp:3039-3063)
			++chainSupport[0][0];
		}
	}

	// path count / DP decision (.cpp:3075-3108)
	int cnt = 0;
	std::vector<int> f((size_t)seCnt, -1);
	bool useDP = false;
	for (i = 0; i < seCnt; ++i)
		if (se[i].canStart)
			cnt = (int)((uint32_t)cnt + (uint32_t)D.subCount(i, f));
	if (cnt <= 200000) {
		for (i = 0; i < seCnt; ++i)
			if (f[i] > 200000) {
				useDP = true;
				break;
			}
	} else
		useDP = true;
	if (!useDP) {
		for (i = 0; i < sampleCnt; ++i) {
			double msize = (double)cs[i].p.size();
			double csize = (double)cs[i].c.size();
			if (cnt > (csize / msize) * (csize / msize) * seCnt && cnt > 200000 / (msize * msize) && cnt > 50) {
				useDP = true;
				break;
			}
		}
	}
	int atCnt = cnt;
	if (stats) {
		stats->atCnt = atCnt;
		stats->useDP = useDP ? 1 : 0;
	}
	std::vector<Transcript> all;
	if (!useDP) {
		std::vector<int> visit((size_t)seCnt + 1);
		for (i = 0; i < seCnt; ++i)
			if (se[i].canStart)
				D.enumerate(i, 0, visit, 0, 1, all);
	} else {
		Decider::DpAttr attr;
		attr.f1.resize((size_t)seCnt);
		attr.hasF2 = seCnt <= 10000;
		if (attr.hasF2) attr.f2.resize((size_t)seCnt * seCnt);
		D.hashMax = 100003;
		if (seCnt > 500) D.hashMax = 1000003; // Q3: the later branches are unreachable
		attr.hash.resize((size_t)D.hashMax);
		std::vector<std::pair<int, int>> complexity((size_t)sampleCnt);
		for (i = 0; i < sampleCnt; ++i)
			complexity[i] = {i, (int)cs[i].c.size()};
		std::stable_sort(complexity.begin(), complexity.end(),
		                 [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second < b.second; }); // Q11
		int downsampleCnt = -1;
		std::vector<Transcript> sampleTx;
		for (i = sampleCnt - 1; i >= 0; --i) {
			sampleTx.clear();
			ConstraintSet &src = cs[complexity[i].first];
			int iterBound = (int)src.c.size();
			if (i < sampleCnt - 1) iterBound = 100;
			if (i < sampleCnt - 10 && all.size() > 1000) {
				iterBound = 10;
				if (stats) ++stats->nIter10;
			}
			if (prm.maxDpConstraintSize > 0) {
				// TruncateConstraintsCoverFrom (Constraints.hpp:400-434)
				std::vector<Constraint> tcv = src.c;
				for (auto &c : tcv) {
					bits_ones(c.v, idx);
					Bits nb = bits_new(seCnt);
					int l = (int)idx.size(), jj;
					for (jj = 0; jj < prm.maxDpConstraintSize && jj < l; ++jj)
						bit_set(nb, idx[jj]);
					c.v = nb;
					c.last = idx[jj - 1];
				}
				D.pickByDP(iterBound, tcv, attr, sampleTx);
			} else if ((src.c.size() > 1000 && src.c.size() * 10 < src.p.size()) ||
			           (downsampleCnt > 0 && (int)src.c.size() >= downsampleCnt) || seCnt >= 1500) {
				int stride = (int)src.p.size() / (int)src.c.size();
				if (downsampleCnt > 0) stride = (int)src.c.size() / downsampleCnt;
				if (stride < 1) stride = 1;
				std::vector<Constraint> tcv; // DownsampleConstraintsFrom (Constraints.hpp:362-398)
				for (size_t q = 0; q < src.c.size(); q += (size_t)stride)
					tcv.push_back(src.c[q]);
				if (downsampleCnt <= 0) downsampleCnt = (int)tcv.size();
				if (stats) ++stats->nDownsampled;
				if (iterBound <= 10) {
					if (stats) ++stats->nSkipped;
					continue;
				}
				D.pickByDP(iterBound, tcv, attr, sampleTx);
			} else
				D.pickByDP(iterBound, src.c, attr, sampleTx);
			for (auto &t : sampleTx) all.push_back(t);
			Decider::coalesceTx(all);
			D.augment(all, 1000);
		}
	}

	// two rounds of per-sample picking (.cpp:3267-3321)
	atCnt = (int)all.size();
	for (auto &t : all) t.FPKM = 0;
	std::vector<std::vector<Transcript>> pred((size_t)sampleCnt);
	for (i = 0; i < sampleCnt; ++i) {
		for (auto &t : all) t.abundance = -1;
		D.pickTranscripts(all, cs[i], pred[i]);
	}
	std::vector<int> txptSampleSupport((size_t)atCnt, 0);
	for (i = 0; i < sampleCnt; ++i)
		for (auto &t : pred[i]) {
			++txptSampleSupport[t.id];
			++all[t.id].FPKM;
		}
	for (i = 0; i < sampleCnt; ++i) {
		for (auto &t : all) t.abundance = -1;
		pred[i].clear();
		D.pickTranscripts(all, cs[i], pred[i]);
	}
	std::vector<std::vector<int>> rawIds((size_t)sampleCnt);
	std::vector<std::vector<double>> rawAbund((size_t)sampleCnt);
	for (i = 0; i < sampleCnt; ++i)
		for (auto &t : pred[i]) {
			rawIds[i].push_back(t.id);
			rawAbund[i].push_back(t.abundance);
		}
	// filtration (.cpp:3336-3353)
	for (i = 0; i < sampleCnt; ++i) {
		for (auto &t : pred[i]) D.toFPKM(t);
		int size = D.refine(false, chainSupport, txptSampleSupport, pred[i], cs[i]);
		D.abundanceEM(cs[i], pred[i]);
		for (j = 0; j < size; ++j) D.toFPKM(pred[i][j]);
		D.refine(true, chainSupport, txptSampleSupport, pred[i], cs[i]);
	}
	// rescue (.cpp:3355-3443)
	std::fill(txptSampleSupport.begin(), txptSampleSupport.end(), 0);
	for (i = 0; i < sampleCnt; ++i)
		for (auto &t : pred[i]) ++txptSampleSupport[t.id];
	std::vector<char> predicted((size_t)atCnt);
	for (i = 0; i < sampleCnt; ++i) {
		std::fill(predicted.begin(), predicted.end(), 0);
		if (pred[i].size() != rawIds[i].size()) {
			int psize = (int)pred[i].size(), rsize = (int)rawIds[i].size();
			for (j = 0; j < psize; ++j) predicted[pred[i][j].id] = 1;
			for (j = 0; j < rsize; ++j) {
				int id = rawIds[i][j];
				if (!predicted[id] && (txptSampleSupport[id] >= 3 && txptSampleSupport[id] >= 0.25 * sampleCnt)) {
					Transcript nt = all[id];
					nt.csupport.clear();
					nt.corr = -1;
					nt.abundance = rawAbund[i][j];
					nt.id = id;
					pred[i].push_back(nt);
				}
			}
			if (psize != (int)pred[i].size())
				D.abundanceEM(cs[i], pred[i]);
		}
		for (auto &t : pred[i])
			D.output(compIdx, comp.chr, i, t, out);
	}
}

} // namespace orc
