// k_dp2.cuh -- SolveSubTranscript / SearchSubTranscript (TranscriptDecider.cpp:374-857) as a
// compiled static program plus an interpreter.  Included by k_dp.cuh.
//
// Observation: the SHAPE of the reference's recursion does not depend on the values it computes.
// Which successors SearchSubTranscript visits (strand rule, .cpp:618-637 / 798-819), which pending
// constraint extensions survive a step (.cpp:407-412), where a walk turns into a new key
// (.cpp:416-422) and which constraints each of the three cover scans (.cpp:463-490, 569-596,
// 732-755) finds compatible are all functions of the graph, the call strand and the constraint bit
// vectors alone.  Only the folds over those scans depend on the threshold (attr.minAbundance) and on
// the constraint abundances, and PickTranscriptsByDP re-runs the recursion hundreds to thousands of
// times per (component, sample) with nothing but those two changing.
//
// So the recursion is compiled once per (key, call strand) into a list of FRAMES, one per visited
// walk position that can produce a candidate, in the reference's depth-first order:
//     nested  : the key SolveSubTranscript would be called on at this position, or none
//     S list  : constraints starting inside the owner key that are compatible with key + walk
//               (what the scan of .cpp:463-490 counts; every such constraint ends inside the walk,
//               because a pending extension reaching further would have postponed the nested call)
//     T list  : S plus the constraints starting inside the walk (the scan of .cpp:569-596)
//     uncov   : the path key + walk crosses an edge no constraint covers (.cpp:438-459)
// with, per list, the static minimum normAbund and the current number of entries whose abundance is
// still positive.  A level of the search then is the reference's own memoised recursion (same call
// order, same hit rule incl. the result-strand comparison, same strict '>' fold), but every scan is
// O(1): "some compatible constraint is at or below the threshold" is minNorm <= threshold, and the
// cover count is the cached number of positive entries, refreshed when an abundance runs dry.
// Keys get exact slots (a table keyed by the key's bit vector) instead of f1 / f2 / a lossy hash.
#pragma once

namespace pscdev {

struct Dp2Key {
	// memo (SetDpContent, .cpp:235-248): valid once cover != -1
	double cover, beta, minAb;
	double ownMinNorm;
	int mlast, mstrand, ts, iter, mode;
	// the key: a set of subexons visited consecutively
	int first, last, cnt, keyUncov, tcStart;
	// own terminal (the key's last subexon can end a transcript): S list of the key alone.
	// ownOff == -2: not built yet, -1: cannot end
	int ownOff, ownCnt, ownPos;
	int prog[3], nfr[3]; // frame list per call strand (-1, 0, +1), prog < 0: not compiled
	int pad;
};

struct Dp2Frame {
	double minNormS, minNormT;
	int nested;  // key id, or -1
	int listOff; // T list = z.list[listOff, listOff + tCnt); its first sCnt entries are the S list
	int sCnt, tCnt;
	int sPos, tPos; // entries with abundance > 0 (dynamic)
	int tag;
	int flags; // bits 0-1: strand + 1; DP2_F_UNCOV; DP2_F_TERM: the tag can end a transcript
};
#define DP2_F_UNCOV 4
#define DP2_F_TERM 8

struct Dp2 {
	Dp2Key *keys;
	uint64_t *kbits, *mbits; // [maxKeys * W] key bits, memo path bits
	int *key1;               // [n] id of the single-subexon key, -1
	int *khash;              // [hashCap] open addressing over kbits, -1 empty
	int hashCap, nKeys, maxKeys;
	Dp2Frame *frames;
	uint64_t *fbits; // [maxFrames * W] key + walk bits of the frame
	int nFrames, maxFrames;
	int *list;
	int nList, maxList;
	// activation stack of the interpreter
	int *aKey, *aS, *aFi, *aEnd, *aLast, *aStrand;
	double *aCover, *aBeta;
	uint64_t *aBits;
	// walk stack of the compiler: index 0 is the key's last subexon
	int *wNode, *wStrand, *wEi, *wExt, *wUncov;
	uint64_t *wBits;
	uint64_t *tmp; // [W]
	int depthCap;
	int overflow;
	size_t hotBytesUsed;
};

// The working set has two parts.  HOT: what the interpreter touches on every step (key records with their memo,
// key / memo / frame bit rows, frame records, the activation stack) -- it lives in the calling group's slice of
// shared memory when it fits there, else in the group's global arena.  COLD: the constraint table handed to the DP,
// the compatibility lists, the compiler's walk stack and the per-threshold result cache -- always global.
#define DP2_FR_PER_KEY_SMEM 3
#define DP2_FR_PER_KEY_GLOBAL 6
#define DP2_FR_PER_KEY_ROOMY 24 // second attempt of a unit that outgrew its arena (long untruncated constraints: ~10 frames per key)
#define DP2_LIST_PER_KEY 64
__host__ __device__ inline size_t dp2_hot_fixed(int n, int W)
{
	size_t d = (size_t)n + 4;
	// 6 int arrays, 2 double arrays, 1 bit-row array (each 16-byte aligned), key1, tmp
	return 6 * ((4 * d + 15) & ~(size_t)15) + 2 * ((8 * d + 15) & ~(size_t)15) + ((8 * d * W + 15) & ~(size_t)15) +
	       ((4 * ((size_t)n + 1) + 15) & ~(size_t)15) + ((8 * (size_t)W + 15) & ~(size_t)15) + 64;
}
__host__ __device__ inline size_t dp2_hot_per_key(int W, int frPerKey)
{
	return sizeof(Dp2Key) + 16 * (size_t)W + 16 + (size_t)frPerKey * (sizeof(Dp2Frame) + 8 * (size_t)W);
}
__host__ __device__ inline int dp2_wanted_keys(int n)
{
	long long k = 8LL * n + 128;
	return (int)(k > 65536 ? 65536 : k);
}
// shared memory a unit is given in a packed launch: room for 2n + 16 keys
__host__ __device__ inline size_t dp2_smem_slice(int n)
{
	int W = (n + 63) >> 6;
	size_t b = dp2_hot_fixed(n, W) + (size_t)(2 * n + 16) * dp2_hot_per_key(W, DP2_FR_PER_KEY_SMEM) + 128;
	return (b + 127) & ~(size_t)127;
}
// shared memory a unit would like: room for 2n + 32 keys (the heaviest units of the bench workload use ~1.4 n)
__host__ __device__ inline size_t dp2_smem_wanted(int n)
{
	int W = (n + 63) >> 6;
	return dp2_hot_fixed(n, W) + (size_t)(2 * n + 32) * dp2_hot_per_key(W, DP2_FR_PER_KEY_SMEM) + 256;
}

// Carves one unit's working set.  cold/coldBytes: the group's global arena; hot/hotBytes: its shared-memory slice or
// nullptr (then the hot part follows the cold part in the arena).  With cold == nullptr and coldBytes == 0 returns the
// arena bytes a unit needs for dp2_wanted_keys(n) keys with everything in global memory.
__host__ __device__ inline size_t dp2_layout(DpScratch *od, Dp2 *oz, char *cold, size_t coldBytes, char *hot, size_t hotBytes, int n,
                                             int E, int Cmax, int keysForSizing = 0, int frPerKeyGlobal = DP2_FR_PER_KEY_GLOBAL)
{
	Carver cv(cold);
	int W = (n + 63) >> 6;
	DpScratch s;
	s.ent = nullptr; s.eepoch = nullptr; s.ebits = nullptr; s.f2Base = 0; s.hashBase = 0; s.hashMax = 0;
	s.P = nullptr; s.fr = nullptr; s.fdcover = nullptr; s.fdbeta = nullptr; s.fdkey = nullptr; s.pbits = nullptr;
	s.fdmeta = nullptr; s.fdbits = nullptr; s.sortidx = nullptr; s.c2b = nullptr; s.c2f = nullptr; s.c2l = nullptr; s.sort64 = nullptr; s.sort64Cap = 0;
	s.txb = nullptr; s.txf = nullptr; s.txl = nullptr; s.txCap = 0;
	s.tcb = cv.take<uint64_t>((size_t)Cmax * W + 1);
	s.tcf = cv.take<int>((size_t)Cmax + 1);
	s.tcl = cv.take<int>((size_t)Cmax + 1);
	s.tcsup = cv.take<int>((size_t)Cmax + 1);
	s.tcnorm = cv.take<double>((size_t)Cmax + 1);
	s.tcab = cv.take<double>((size_t)Cmax + 1);
	s.tcCnt = 0;
	s.uncov = cv.take<uint8_t>((size_t)E + 1);
	s.tcEnd = cv.take<int>((size_t)n + 2);
	s.extCap = 2 * (Cmax + 1) + 64;
	s.ext = cv.take<int>((size_t)s.extCap);
	s.cCap = Cmax + 2;
	s.ckey = cv.take<double>((size_t)s.cCap);
	s.cent = cv.take<DpMemo>((size_t)s.cCap);
	s.cbits = cv.take<uint64_t>((size_t)s.cCap * W);
	s.cvalid = cv.take<uint8_t>((size_t)s.cCap);
	s.work = cv.take<uint64_t>((size_t)8 * W);
	s.covered = cv.take<int>((size_t)Cmax + 1);
	Dp2 z;
	z.depthCap = n + 4;
	z.wNode = cv.take<int>((size_t)z.depthCap);
	z.wStrand = cv.take<int>((size_t)z.depthCap);
	z.wEi = cv.take<int>((size_t)z.depthCap);
	z.wExt = cv.take<int>((size_t)z.depthCap);
	z.wUncov = cv.take<int>((size_t)z.depthCap);
	z.wBits = cv.take<uint64_t>((size_t)z.depthCap * W);
	const size_t coldFixed = (cv.off + 15) & ~(size_t)15;
	const size_t hotFixed = dp2_hot_fixed(n, W);
	const bool split = hot != nullptr;
	const int frPerKey = split ? DP2_FR_PER_KEY_SMEM : frPerKeyGlobal;
	long long K = keysForSizing > 0 ? keysForSizing : dp2_wanted_keys(n), L;
	if (coldBytes > 0) {
		if (split) {
			K = hotBytes > hotFixed + 64 ? (long long)((hotBytes - hotFixed - 64) / dp2_hot_per_key(W, frPerKey)) : 0;
			L = coldBytes > coldFixed + 64 ? (long long)((coldBytes - coldFixed - 64) / sizeof(int)) : 0;
			if (K > dp2_wanted_keys(n)) K = dp2_wanted_keys(n);
		} else {
			size_t perKey = dp2_hot_per_key(W, frPerKey) + DP2_LIST_PER_KEY * sizeof(int);
			K = coldBytes > coldFixed + hotFixed + 64 ? (long long)((coldBytes - coldFixed - hotFixed - 64) / perKey) : 0;
			L = DP2_LIST_PER_KEY * K;
		}
		if (K > 1 << 20) K = 1 << 20;
	} else
		L = DP2_LIST_PER_KEY * K;
	if (L > (long long)DP2_LIST_PER_KEY * K) L = (long long)DP2_LIST_PER_KEY * K;
	z.maxKeys = (int)K;
	z.maxFrames = (int)(frPerKey * K);
	z.maxList = (int)L;
	z.hashCap = 4;
	while (z.hashCap < 2 * K) z.hashCap <<= 1; // < 4K slots
	z.list = cv.take<int>((size_t)z.maxList);
	Carver hv(split ? hot : cold);
	if (!split) hv.off = (cv.off + 15) & ~(size_t)15;
	z.aKey = hv.take<int>((size_t)z.depthCap);
	z.aS = hv.take<int>((size_t)z.depthCap);
	z.aFi = hv.take<int>((size_t)z.depthCap);
	z.aEnd = hv.take<int>((size_t)z.depthCap);
	z.aLast = hv.take<int>((size_t)z.depthCap);
	z.aStrand = hv.take<int>((size_t)z.depthCap);
	z.aCover = hv.take<double>((size_t)z.depthCap);
	z.aBeta = hv.take<double>((size_t)z.depthCap);
	z.aBits = hv.take<uint64_t>((size_t)z.depthCap * W);
	z.tmp = hv.take<uint64_t>((size_t)W);
	z.key1 = hv.take<int>((size_t)n + 1);
	z.keys = hv.take<Dp2Key>((size_t)z.maxKeys);
	z.kbits = hv.take<uint64_t>((size_t)z.maxKeys * W);
	z.mbits = hv.take<uint64_t>((size_t)z.maxKeys * W);
	z.khash = hv.take<int>((size_t)z.hashCap);
	z.frames = hv.take<Dp2Frame>((size_t)z.maxFrames);
	z.fbits = hv.take<uint64_t>((size_t)z.maxFrames * W);
	z.hotBytesUsed = split ? hv.off : 0;
	z.nKeys = z.nFrames = z.nList = 0;
	z.overflow = 0;
	if (od) *od = s;
	if (oz) *oz = z;
	return split ? ((cv.off + 15) & ~(size_t)15) : ((hv.off + 15) & ~(size_t)15);
}

PSC_DEV uint32_t dp2_hash_bits(const uint64_t *b, int W)
{
	uint64_t h = 0x9E3779B97F4A7C15ull;
	for (int w = 0; w < W; ++w) {
		h ^= b[w] + (uint64_t)w;
		h *= 0xff51afd7ed558ccdull;
		h ^= h >> 29;
	}
	return (uint32_t)(h ^ (h >> 32));
}

// id of the key with these bits, created on first sight.  Every lane executes this with the same
// arguments and writes the same values, so no lane election or synchronisation is needed.
PSC_DEV int dp2_key(DpState &st, const uint64_t *bits, int first, int last, int cnt, int keyUncov)
{
	Dp2 &z = *st.z;
	const int W = st.W;
	int *slot;
	if (cnt == 1) {
		slot = z.key1 + last;
		if (*slot >= 0) return *slot;
	} else {
		uint32_t h = dp2_hash_bits(bits, W) & (uint32_t)(z.hashCap - 1);
		for (;;) {
			int id = z.khash[h];
			if (id < 0) break;
			const uint64_t *kb = z.kbits + (long long)id * W;
			bool same = true;
			for (int w = 0; w < W; ++w) same = same && kb[w] == bits[w];
			if (same) return id;
			h = (h + 1) & (uint32_t)(z.hashCap - 1);
		}
		slot = z.khash + h;
	}
	if (z.nKeys >= z.maxKeys) {
		z.overflow = 1;
		return -1;
	}
	int id = z.nKeys++;
	Dp2Key *k = z.keys + id;
	k->cover = -1;
	k->beta = 1e300;
	k->minAb = 0;
	k->ownMinNorm = 1e300;
	k->mlast = last;
	k->mstrand = 0;
	k->ts = 0;
	k->iter = 0;
	k->mode = 0;
	k->first = first;
	k->last = last;
	k->cnt = cnt;
	k->keyUncov = keyUncov;
	k->tcStart = first > 0 ? st.d.tcEnd[first - 1] : 0;
	k->ownOff = -2;
	k->ownCnt = 0;
	k->ownPos = 0;
	for (int q = 0; q < 3; ++q) {
		k->prog[q] = -1;
		k->nfr[q] = 0;
	}
	uint64_t *kb = z.kbits + (long long)id * W;
	for (int w = 0; w < W; ++w) kb[w] = bits[w];
	*slot = id;
	return id;
}

struct Dp2List {
	int off, cnt, sCnt, pos, sPos;
	double mn, smn;
};
// constraints [from, to) compatible with the complete transcript (tb, tf, tl), appended to z.list in index order;
// the entries below sTo form the S part
PSC_DEV bool dp2_build_list(DpState &st, const uint64_t *tb, int tf, int tl, int from, int to, int sTo, Dp2List &r)
{
	const Coop &co = st.co;
	const DpScratch &d = st.d;
	Dp2 &z = *st.z;
	const int W = st.W;
	r.off = z.nList;
	int cnt = 0, myS = 0, myPos = 0, mySPos = 0;
	double mn = 1e300, smn = 1e300;
	for (int b0 = from; b0 < to; b0 += co.n) {
		int i = b0 + co.lane;
		bool is = i < to && compat(tb, tf, tl, false, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 1;
		uint32_t m = co.ballot(is);
		if (is) {
			int p = z.nList + cnt + co.rank_in(m);
			if (p < z.maxList) z.list[p] = i;
			double nm = d.tcnorm[i];
			bool pos = d.tcab[i] > 0;
			if (nm < mn) mn = nm;
			myPos += pos ? 1 : 0;
			if (i < sTo) {
				++myS;
				mySPos += pos ? 1 : 0;
				if (nm < smn) smn = nm;
			}
		}
		cnt += popc32(m);
	}
	if (z.nList + cnt > z.maxList) {
		z.overflow = 1;
		return false;
	}
	z.nList += cnt;
	r.cnt = cnt;
	r.sCnt = co.sum(myS);
	r.pos = co.sum(myPos);
	r.sPos = co.sum(mySPos);
	r.mn = co.min(mn);
	r.smn = co.min(smn);
	co.sync();
	return true;
}

// Compiles SolveSubTranscript(key K, call strand si - 1) into frames.  No values are involved.
PSC_DEVN void dp2_compile(DpState &st, int K, int si)
{
	const Coop co = st.co;
	const CompView &v = st.v;
	const DpScratch &d = st.d;
	Dp2 &z = *st.z;
	const int W = st.W;
	Dp2Key *key = z.keys + K;
	const uint64_t *kb = z.kbits + (long long)K * W;
	const int kfirst = key->first, klast = key->last, tcStart = key->tcStart;
	const int sEnd = dp_tc_end(st, tcStart, klast); // constraints starting inside the key: [tcStart, sEnd)
	const int e0 = v.noff[0];
	if (key->ownOff == -2) {
		if (st.ce[klast]) {
			Dp2List r;
			if (!dp2_build_list(st, kb, kfirst, klast, tcStart, sEnd, sEnd, r)) return;
			key->ownOff = r.off;
			key->ownCnt = r.cnt;
			key->ownPos = r.pos;
			key->ownMinNorm = r.mn;
		} else
			key->ownOff = -1;
	}
	// extensions: constraints whose head matches the key and that reach beyond it (.cpp:765-796)
	int extCnt = 0;
	for (int b0 = tcStart; b0 < sEnd; b0 += co.n) {
		int i = b0 + co.lane;
		bool is = i < sEnd && compat(kb, kfirst, klast, true, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 2;
		uint32_t m = co.ballot(is);
		if (is) {
			int p = extCnt + co.rank_in(m);
			if (p < d.extCap) d.ext[p] = i;
		}
		extCnt += popc32(m);
	}
	if (2 * extCnt > d.extCap) { // cannot happen: extCap = 2 * (C + 1) + 64
		z.overflow = 1;
		return;
	}
	co.sync();
	dp_sort_ext(st, d.ext, extCnt, true);
	const int base = z.nFrames;
	int nf = 0;
	int p = 0;
	z.wNode[0] = klast;
	z.wStrand[0] = si - 1;
	z.wEi[0] = v.noff[klast];
	z.wExt[0] = extCnt;
	z.wUncov[0] = key->keyUncov;
	for (int w = 0; w < W; ++w) z.wBits[w] = kb[w];
	for (;;) {
		const int a = z.wNode[p], cs = z.wStrand[p];
		int ei = z.wEi[p];
		int b = -1, ns = cs;
		{
			const int eEnd = v.noff[a + 1], rsa = v.rs[a], ena = v.en[a];
			const bool aOk = same_strand(rsa, cs);
			while (ei < eEnd) {
				int x = v.nidx[ei++];
				int stx = v.st[x];
				if ((aOk && same_strand(v.ls[x], cs)) || stx == ena + 1) {
					b = x;
					if (stx > ena + 1) ns = rsa;
					break;
				}
			}
		}
		z.wEi[p] = ei;
		if (b < 0) {
			if (p == 0) break;
			--p;
			continue;
		}
		const int np = p + 1;
		if (np >= z.depthCap) {
			z.overflow = 1;
			return;
		}
		const int nuncov = z.wUncov[p] + (d.uncov[ei - 1 - e0] ? 1 : 0);
		uint64_t *nb = z.wBits + (long long)np * W;
		{
			const uint64_t *pb = z.wBits + (long long)p * W;
			for (int w = 0; w < W; ++w) nb[w] = pb[w];
			bt_set(nb, b);
		}
		z.wNode[np] = b;
		z.wStrand[np] = ns;
		z.wUncov[np] = nuncov;
		// drop the pending extensions that no longer fit (.cpp:407-412)
		int keptLast = 0, i;
		for (i = z.wExt[p] - 1; i >= 0; --i) {
			int x = d.ext[i];
			int xl = d.tcl[x];
			keptLast = xl;
			if (xl <= b) break;
			if (bt_test(d.tcb + (long long)x * W, b) && compat(nb, kfirst, b, true, d.tcb + (long long)x * W, d.tcf[x], xl, W) != 0) break;
		}
		const int next = i + 1;
		z.wExt[np] = next;
		const bool solveHere = v.noff[b + 1] > v.noff[b] && (next == 0 || b >= keptLast);
		const bool canEnd = st.ce[b] != 0;
		if (solveHere || canEnd) {
			if (z.nFrames >= z.maxFrames) {
				z.overflow = 1;
				return;
			}
			int nested = -1;
			if (solveHere) {
				for (int w = 0; w < W; ++w) z.tmp[w] = nb[w] & ~kb[w];
				nested = dp2_key(st, z.tmp, z.wNode[1], b, np, np == 1 ? 0 : nuncov - z.wUncov[1]);
				if (nested < 0) return;
				key = z.keys + K;
			}
			Dp2List r;
			if (!dp2_build_list(st, nb, kfirst, b, tcStart, canEnd ? dp_tc_end(st, tcStart, b) : sEnd, sEnd, r)) return;
			const int f = z.nFrames++;
			Dp2Frame *fr = z.frames + f;
			fr->minNormS = r.smn;
			fr->minNormT = r.mn;
			fr->nested = nested;
			fr->listOff = r.off;
			fr->sCnt = r.sCnt;
			fr->tCnt = r.cnt;
			fr->sPos = r.sPos;
			fr->tPos = r.pos;
			fr->tag = b;
			fr->flags = (ns + 1) | (nuncov > 0 ? DP2_F_UNCOV : 0) | (canEnd ? DP2_F_TERM : 0);
			uint64_t *fb = z.fbits + (long long)f * W;
			for (int w = 0; w < W; ++w) fb[w] = nb[w];
			++nf;
		}
		if (!solveHere) {
			p = np;
			z.wEi[np] = v.noff[b];
		}
	}
	key->prog[si] = base;
	key->nfr[si] = nf;
	co.sync();
}

// the three cover scans with everything static folded into (cnt, pos, minNorm)
PSC_DEV double dp2_scan(const DpState &st, const Dp2 &z, int off, int cnt, int pos, double minNorm, double cover, bool &below,
                        double &beta)
{
	if (cnt <= 0) return cover;
	if (minNorm <= st.minAbundance) { // the reference stops at the first such constraint and reports -2
		below = true;
		return -2;
	}
	if (!st.forAbundance) {
		if (minNorm < beta) beta = minNorm;
		return cover + pos;
	}
	if (pos > 0) {
		double mn = minNorm;
		if (pos != cnt) { // some entries are not positive (cannot happen before the first subtraction; kept exact)
			mn = 1e300;
			for (int q = st.co.lane; q < cnt; q += st.co.n) {
				int i = z.list[off + q];
				if (st.d.tcab[i] > 0 && st.d.tcnorm[i] < mn) mn = st.d.tcnorm[i];
			}
			mn = st.co.min(mn);
		}
		if (cover == 0 || mn < cover) cover = mn;
	}
	return cover;
}

PSC_DEV bool dp2_hit(const DpState &st, const Dp2Key *k, int strand)
{
	if (k->cover == -1 || k->mstrand != strand) return false;
	if (k->ts == st.timeStamp) return true;
	if (k->minAb < st.minAbundance && k->cover == -2) return true;
	// see dp_memo_hit
	return !st.forAbundance && k->mode == 0 && k->iter == st.iter && k->cover >= 0 && k->minAb > 0 && st.minAbundance > k->minAb &&
	       k->beta > st.minAbundance;
}

// SolveSubTranscript({start}, strand 0): returns the key id whose memo holds the result, -1 on overflow
PSC_DEVN int dp2_solve_root(DpState &stRef, int start)
{
	Dp2 &z = *stRef.z;
	const Coop co = stRef.co;
	const int W = stRef.W;
	for (int w = 0; w < W; ++w) z.tmp[w] = 0;
	bt_set(z.tmp, start);
	const int root = dp2_key(stRef, z.tmp, start, start, 1, 0);
	if (root < 0) return -1;
	const bool forAb = stRef.forAbundance;
	const double tau = stRef.minAbundance;
	int depth = -1;
	int nStates = 0, nFrames = 0;
	// the top activation
	int t_key = 0, t_fi = 0, t_base = 0, t_nfr = 0;
	double o_cover = -1, o_beta = 1e300;
	int o_plast = 0, o_pstrand = 0;
	const uint64_t *t_kb = nullptr;
	uint64_t *t_bits = nullptr;
	int rk = root, rs = 0, ret = -1;
	int mode = 0; // 0 call (rk, rs), 1 next frame, 2 evaluate frame t_fi with the sub-solution ret
	for (;;) {
		if (mode == 0) {
			Dp2Key *k = z.keys + rk;
			if (dp2_hit(stRef, k, rs)) {
				ret = rk;
				if (depth < 0) break;
				mode = 2;
				continue;
			}
			if (k->prog[rs + 1] < 0) {
				dp2_compile(stRef, rk, rs + 1);
				if (z.overflow) return -1;
				k = z.keys + rk;
			}
			++nStates;
			if (depth >= 0) {
				z.aKey[depth] = t_key;
				z.aFi[depth] = t_fi;
				z.aCover[depth] = o_cover;
				z.aBeta[depth] = o_beta;
				z.aLast[depth] = o_plast;
				z.aStrand[depth] = o_pstrand;
			}
			++depth;
			if (depth >= z.depthCap) {
				z.overflow = 1;
				return -1;
			}
			z.aS[depth] = rs;
			t_key = rk;
			t_fi = 0;
			t_base = k->prog[rs + 1];
			t_nfr = k->nfr[rs + 1];
			t_kb = z.kbits + (long long)rk * W;
			t_bits = z.aBits + (long long)depth * W;
			// pdp starts as the key itself (.cpp:722-760)
			o_cover = -1;
			o_beta = 1e300;
			o_plast = k->last;
			o_pstrand = rs;
			for (int w = 0; w < W; ++w) t_bits[w] = t_kb[w];
			if (k->ownOff >= 0) {
				bool bm = false;
				o_cover = dp2_scan(stRef, z, k->ownOff, k->ownCnt, k->ownPos, k->ownMinNorm, 0, bm, o_beta);
			}
			mode = 1;
			continue;
		}
		if (mode == 1) {
			if (t_fi == t_nfr) {
				// the activation is finished: SetDpContent (.cpp:824-856)
				Dp2Key *k = z.keys + t_key;
				uint64_t *mb = z.mbits + (long long)t_key * W;
				for (int w = 0; w < W; ++w) mb[w] = t_bits[w];
				k->mlast = o_plast;
				k->mstrand = o_pstrand;
				k->cover = o_cover;
				k->beta = o_beta;
				k->iter = stRef.iter;
				k->mode = forAb ? 1 : 0;
				k->minAb = tau;
				k->ts = stRef.timeStamp;
				ret = t_key;
				--depth;
				if (depth < 0) break;
				t_key = z.aKey[depth];
				t_fi = z.aFi[depth];
				o_cover = z.aCover[depth];
				o_beta = z.aBeta[depth];
				o_plast = z.aLast[depth];
				o_pstrand = z.aStrand[depth];
				{
					const Dp2Key *pk = z.keys + t_key;
					const int ps = z.aS[depth];
					t_base = pk->prog[ps + 1];
					t_nfr = pk->nfr[ps + 1];
				}
				t_kb = z.kbits + (long long)t_key * W;
				t_bits = z.aBits + (long long)depth * W;
				mode = 2;
				continue;
			}
			const Dp2Frame *fr = z.frames + t_base + t_fi;
			if (fr->nested >= 0) {
				rk = fr->nested;
				rs = (fr->flags & 3) - 1;
				mode = 0;
				continue;
			}
			ret = -1;
			mode = 2;
			continue;
		}
		// ---- evaluate frame t_fi given the sub-solution (.cpp:425-614) ----
		{
			++nFrames;
			const Dp2Frame fr = z.frames[t_base + t_fi];
			const int fstrand = (fr.flags & 3) - 1;
			const bool uncov = (fr.flags & DP2_F_UNCOV) != 0;
			double vcover = -1, vbeta = 1e300;
			int vlast = -1;
			const uint64_t *vbits = nullptr;
			if (fr.nested >= 0) {
				const Dp2Key *vk = z.keys + ret;
				vcover = vk->cover;
				vbeta = vk->beta;
				vlast = vk->mlast;
				vbits = z.mbits + (long long)ret * W;
			}
			if (vcover >= 0) {
				double cover = vcover;
				bool belowMin = !forAb && tau > 0 && uncov;
				bool bm2 = false;
				double cbeta = vbeta;
				cover = dp2_scan(stRef, z, fr.listOff, fr.sCnt, fr.sPos, fr.minNormS, cover, bm2, cbeta);
				if (bm2) belowMin = true;
				bool takeBelow = belowMin && o_cover == -1;
				if (takeBelow || (!(belowMin && o_cover == -1) && cover > o_cover)) {
					o_cover = takeBelow ? -2 : cover;
					o_beta = cbeta;
					o_plast = vlast;
					o_pstrand = fstrand;
					for (int w = 0; w < W; ++w) t_bits[w] = t_kb[w] | vbits[w];
				}
			} else if (vcover == -2 && o_cover == -1) {
				o_cover = -2;
				o_plast = vlast;
				o_pstrand = fstrand;
				for (int w = 0; w < W; ++w) t_bits[w] = t_kb[w] | vbits[w];
			}
			// the path may also stop at this subexon (.cpp:524-614)
			if ((fr.flags & DP2_F_TERM) && (vcover < 0 || forAb)) {
				bool below = false;
				double cover = 0;
				if (forAb || tau > 0) {
					if (uncov) below = true;
					if (forAb && below) cover = 1e-6;
				}
				bool bm2 = false;
				double cbeta = 1e300;
				cover = dp2_scan(stRef, z, fr.listOff, fr.tCnt, fr.tPos, fr.minNormT, cover, bm2, cbeta);
				if (bm2) below = true;
				bool takeBelow = below && o_cover == -1;
				if (takeBelow || (!(below && o_cover == -1) && cover > o_cover)) {
					o_cover = takeBelow ? -2 : cover;
					o_beta = cbeta;
					o_plast = fr.tag;
					o_pstrand = fstrand;
					const uint64_t *fb = z.fbits + (long long)(t_base + t_fi) * W;
					for (int w = 0; w < W; ++w) t_bits[w] = fb[w];
				}
			}
			++t_fi;
			mode = 1;
		}
	}
	co.sync();
	stRef.nStates += nStates;
	stRef.nFrames += nFrames;
	++stRef.nRoots;
	return ret;
}

// The same interpreter for the common case (cover counting, forAbundance == false), NW = words per bit row known at
// compile time: every table base pointer, the threshold and the current activation (best cover / beta / path bits
// included) are registers, the three modes are straight-line code joined by gotos, and the cover scans reduce to a
// comparison and an addition.  Results are identical to dp2_solve_root's by construction (same order, same rules).
template <int NW> PSC_DEVN int dp3_solve_root(DpState &stRef, int start)
{
	Dp2 &z = *stRef.z;
	if (z.key1[start] < 0) {
		for (int w = 0; w < NW; ++w) z.tmp[w] = 0;
		bt_set(z.tmp, start);
		if (dp2_key(stRef, z.tmp, start, start, 1, 0) < 0) return -1;
	}
	Dp2Key *const keys = z.keys;
	const Dp2Frame *const frames = z.frames;
	const uint64_t *const kbits = z.kbits, *const fbits = z.fbits;
	uint64_t *const mbits = z.mbits, *const aBits = z.aBits;
	int *const aKey = z.aKey, *const aFi = z.aFi, *const aEnd = z.aEnd, *const aLast = z.aLast, *const aStrand = z.aStrand;
	double *const aCover = z.aCover, *const aBeta = z.aBeta;
	const int depthCap = z.depthCap;
	const double tau = stRef.minAbundance;
	const bool tauPos = tau > 0;
	const int ts = stRef.timeStamp, iter = stRef.iter;
	int depth = -1, nStates = 0, nFrames = 0;
	int t_key = 0, t_fi = 0, t_end = 0;
	double o_cover = -1, o_beta = 1e300;
	int o_plast = 0, o_pstrand = 0;
	uint64_t ob[NW];
	for (int w = 0; w < NW; ++w) ob[w] = 0;
	int rk = z.key1[start], rs = 0, ret = -1;
call : {
	Dp2Key *k = keys + rk;
	{
		// dp2_hit with forAbundance == false
		const double kc = k->cover;
		bool hit = false;
		if (kc != -1 && k->mstrand == rs) {
			const double kmin = k->minAb;
			hit = k->ts == ts || (kmin < tau && kc == -2) ||
			      (k->mode == 0 && k->iter == iter && kc >= 0 && kmin > 0 && tau > kmin && k->beta > tau);
		}
		if (hit) {
			ret = rk;
			if (depth < 0) goto done;
			goto eval;
		}
	}
	if (k->prog[rs + 1] < 0) {
		dp2_compile(stRef, rk, rs + 1);
		if (z.overflow) return -1;
	}
	++nStates;
	if (depth >= 0) {
		aKey[depth] = t_key;
		aFi[depth] = t_fi;
		aEnd[depth] = t_end;
		aCover[depth] = o_cover;
		aBeta[depth] = o_beta;
		aLast[depth] = o_plast;
		aStrand[depth] = o_pstrand;
#pragma unroll
		for (int w = 0; w < NW; ++w) aBits[depth * NW + w] = ob[w];
	}
	++depth;
	if (depth >= depthCap) {
		z.overflow = 1;
		return -1;
	}
	t_key = rk;
	t_fi = k->prog[rs + 1];
	t_end = t_fi + k->nfr[rs + 1];
	// pdp starts as the key itself (.cpp:722-760)
	o_cover = -1;
	o_beta = 1e300;
	o_plast = k->last;
	o_pstrand = rs;
#pragma unroll
	for (int w = 0; w < NW; ++w) ob[w] = kbits[rk * NW + w];
	if (k->ownOff >= 0) {
		o_cover = 0;
		if (k->ownCnt > 0) {
			const double mn = k->ownMinNorm;
			if (mn <= tau) o_cover = -2;
			else {
				o_beta = mn;
				o_cover = k->ownPos;
			}
		}
	}
}
next:
	if (t_fi == t_end) {
		// the activation is finished: SetDpContent (.cpp:824-856)
		Dp2Key *k = keys + t_key;
#pragma unroll
		for (int w = 0; w < NW; ++w) mbits[t_key * NW + w] = ob[w];
		k->mlast = o_plast;
		k->mstrand = o_pstrand;
		k->cover = o_cover;
		k->beta = o_beta;
		k->iter = iter;
		k->mode = 0;
		k->minAb = tau;
		k->ts = ts;
		ret = t_key;
		--depth;
		if (depth < 0) goto done;
		t_key = aKey[depth];
		t_fi = aFi[depth];
		t_end = aEnd[depth];
		o_cover = aCover[depth];
		o_beta = aBeta[depth];
		o_plast = aLast[depth];
		o_pstrand = aStrand[depth];
#pragma unroll
		for (int w = 0; w < NW; ++w) ob[w] = aBits[depth * NW + w];
		goto eval;
	}
	{
		const int nested = frames[t_fi].nested;
		if (nested >= 0) {
			rk = nested;
			rs = (frames[t_fi].flags & 3) - 1;
			goto call;
		}
		ret = -1;
	}
eval : {
	// ---- evaluate frame t_fi given the sub-solution `ret` (.cpp:425-614) ----
	++nFrames;
	const Dp2Frame *fr = frames + t_fi;
	const int flags = fr->flags;
	const int fstrand = (flags & 3) - 1;
	const bool uncov = (flags & DP2_F_UNCOV) != 0;
	double vcover = -1;
	if (ret >= 0) {
		const Dp2Key *vk = keys + ret;
		vcover = vk->cover;
		if (vcover >= 0) {
			double cover = vcover, cbeta = vk->beta;
			bool belowMin = tauPos && uncov;
			if (fr->sCnt > 0) {
				const double mn = fr->minNormS;
				if (mn <= tau) {
					belowMin = true;
					cover = -2;
				} else {
					if (mn < cbeta) cbeta = mn;
					cover += fr->sPos;
				}
			}
			const bool first = belowMin && o_cover == -1;
			if (first || (!first && cover > o_cover)) {
				o_cover = first ? -2 : cover;
				o_beta = cbeta;
				o_plast = vk->mlast;
				o_pstrand = fstrand;
#pragma unroll
				for (int w = 0; w < NW; ++w) ob[w] = kbits[t_key * NW + w] | mbits[ret * NW + w];
			}
		} else if (vcover == -2 && o_cover == -1) {
			o_cover = -2;
			o_plast = vk->mlast;
			o_pstrand = fstrand;
#pragma unroll
			for (int w = 0; w < NW; ++w) ob[w] = kbits[t_key * NW + w] | mbits[ret * NW + w];
		}
	}
	// the path may also stop at this subexon (.cpp:524-614)
	if ((flags & DP2_F_TERM) && vcover < 0) {
		bool below = tauPos && uncov;
		double cover = 0, cbeta = 1e300;
		if (fr->tCnt > 0) {
			const double mn = fr->minNormT;
			if (mn <= tau) {
				below = true;
				cover = -2;
			} else {
				cbeta = mn;
				cover = fr->tPos;
			}
		}
		const bool first = below && o_cover == -1;
		if (first || (!first && cover > o_cover)) {
			o_cover = first ? -2 : cover;
			o_beta = cbeta;
			o_plast = fr->tag;
			o_pstrand = fstrand;
#pragma unroll
			for (int w = 0; w < NW; ++w) ob[w] = fbits[t_fi * NW + w];
		}
	}
	++t_fi;
	goto next;
}
done:
	stRef.nStates += nStates;
	stRef.nFrames += nFrames;
	++stRef.nRoots;
	return ret;
}

// an abundance ran dry: recount the positive entries of every list
PSC_DEV void dp2_refresh(DpState &st)
{
	const Coop &co = st.co;
	Dp2 &z = *st.z;
	const double *ab = st.d.tcab;
	co.sync();
	for (int f = co.lane; f < z.nFrames; f += co.n) {
		Dp2Frame *fr = z.frames + f;
		int sp = 0, tp = 0;
		const int *l = z.list + fr->listOff;
		for (int q = 0; q < fr->tCnt; ++q) {
			int pos = ab[l[q]] > 0 ? 1 : 0;
			tp += pos;
			if (q < fr->sCnt) sp += pos;
		}
		fr->sPos = sp;
		fr->tPos = tp;
	}
	for (int k = co.lane; k < z.nKeys; k += co.n) {
		Dp2Key *key = z.keys + k;
		if (key->ownOff < 0) continue;
		int p = 0;
		const int *l = z.list + key->ownOff;
		for (int q = 0; q < key->ownCnt; ++q) p += ab[l[q]] > 0 ? 1 : 0;
		key->ownPos = p;
	}
	co.sync();
}

struct DpEngineProgram {
	static PSC_DEV DpRootRes root(DpState &st, int i)
	{
		DpRootRes o;
		o.cover = -1;
		o.bits = st.z->mbits;
		o.first = o.last = i;
		if (st.z->overflow) return o;
		int r;
		if (st.forAbundance) r = dp2_solve_root(st, i);
		else if (st.W == 1) r = dp3_solve_root<1>(st, i);
		else if (st.W == 2) r = dp3_solve_root<2>(st, i);
		else if (st.W == 3) r = dp3_solve_root<3>(st, i);
		else if (st.W == 4) r = dp3_solve_root<4>(st, i);
		else r = dp2_solve_root(st, i);
		if (r < 0) return o;
		const Dp2Key *k = st.z->keys + r;
		o.cover = k->cover;
		o.bits = st.z->mbits + (long long)r * st.W;
		o.first = k->first;
		o.last = k->mlast;
		return o;
	}
	static PSC_DEV void abundances_changed(DpState &st) { dp2_refresh(st); }
};

} // namespace pscdev
