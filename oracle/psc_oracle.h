// psc_oracle -- CPU RESTATEMENT of PsiCLASS stage 3 (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// Serial, straightforward C++ restating the reference's algorithm for the hot path
// (Constraints.cpp, TranscriptDecider.cpp; every function cites the lines it follows) on the
// flat structure-of-arrays inputs of the C ABI (include/psc.h).  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may call it; the product
// (psiclass_b200/) never links, imports or executes anything under oracle/.
//
// Pinning: there are no golden vectors in the reference (it has no tests).  This restatement is
// pinned against (a) SURVEY.md Appendix A's known-answer vector for example/ and (b) the
// unmodified reference built into oracle/_ref (classes / ref_dump) on seeded synthetic sets:
// tests/test_oracle_vs_ref.py, fixtures under tests/golden/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace orc {

typedef std::vector<uint64_t> Bits;

struct Subexon {
	int start, end;
	int lt, rt, ls, rs;
	int geneId;
	double lcls, rcls;
	std::vector<int> next, prev; // local indices
	bool canStart = false, canEnd = false;
};

struct Component {
	int chr = 0;
	int baseGeneId = 0, usedGeneId = 0;
	std::vector<Subexon> se;
};

// one sample's alignments (the reads Alignments::Next() returns), structure of arrays
struct Reads {
	std::vector<int64_t> chr_off;
	const int32_t *pos = nullptr, *mate_pos = nullptr;
	const uint8_t *flags = nullptr, *nh = nullptr;
	const uint64_t *name_key = nullptr;
	const uint32_t *seg_off = nullptr;
	const int32_t *seg_a = nullptr, *seg_b = nullptr;
	int64_t n = 0;
	int read_len = 0, frag_len = 0, frag_stdev = 0, mate_paired = 0;
	int64_t cursor = 0;     // stream position (BuildConstraints consumes the BAM forward only)
	bool held = false;      // the read at `cursor` was examined by the previous component and kept
	int64_t last_read = -1; // index of the last read the stream delivered (Q8: totalReadCnt)
};

struct Constraint {
	Bits v;
	double weight = 0, normAbund = 0, abundance = 0;
	int support = 0, uniq = 0, maxReadLen = 0, info = 0, first = 0, last = 0;
};
struct MatePair {
	int i = 0, j = 0, support = 0, uniq = 0;
	double abundance = 0, normAbund = 0;
	int effectiveCount = 0;
};
struct ConstraintSet {
	std::vector<Constraint> c;
	std::vector<MatePair> p;
};

struct Params {
	double fpkmFraction = 0.05, classifierThreshold = 0.05, minDepth = 2.5;
	int maxDpConstraintSize = 7;
	bool primaryAsUnique = false;
	int sampleCnt = 1;
	// from BAM #0 (TranscriptDecider holds alignmentFiles[0]; classes.cpp:287)
	int readLen = 0, fragLen = 0, fragStdev = 0;
	bool matePaired = false;
};

struct OutTranscript {
	int comp = 0, sample = 0, chr = 0, geneId = 0;
	char strand = '.';
	std::vector<std::pair<int, int>> exons; // 1-based inclusive
	double fpkm = 0, cov = 0;               // fpkm is the raw value before ComputeFPKMTPM
};

struct SolveStats {
	int atCnt = 0;
	int useDP = 0;
	// which branches of the DP sample loop (TranscriptDecider.cpp:3196-3244) the component took (test coverage evidence)
	int nDownsampled = 0; // samples that went through DownsampleConstraintsFrom
	int nIter10 = 0;      // samples that got iterBound = 10
	int nSkipped = 0;     // down-sampled samples skipped because iterBound <= 10
};

// Constraints::BuildConstraints (Constraints.cpp:240-438) for one sample and one component;
// advances r.cursor exactly like the reference's forward-only stream.
void build_constraints(const Component &comp, Reads &r, bool primaryAsUnique, ConstraintSet &out);

// TranscriptDecider::Solve (TranscriptDecider.cpp:2874-3460); appends to `out` in the reference's
// queue order (sample-major inside the component).
void solve(int compIdx, Component &comp, std::vector<ConstraintSet> &cs, const Params &prm,
           std::vector<OutTranscript> &out, SolveStats *stats);

// helpers shared with the tests
inline int nwords(int n) { return n <= 0 ? 1 : (n + 63) / 64; }
inline bool bit_test(const Bits &b, int i) { return (b[i >> 6] >> (i & 63)) & 1; }
inline void bit_set(Bits &b, int i) { b[i >> 6] |= (uint64_t)1 << (i & 63); }

} // namespace orc
