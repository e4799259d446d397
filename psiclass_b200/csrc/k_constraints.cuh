// k_constraints.cuh -- alignment -> subexon bit-vector constraints for one (sample, component).
//
// Replaces Constraints::BuildConstraints and its helpers:
//   read -> component assignment   Constraints.cpp:259-283  (closed form, see group_ranges)
//   ConvertAlignmentToBitTable     Constraints.cpp:4-72
//   clip rule                      Constraints.cpp:310-326
//   MateReadIds (mate join)        Constraints.hpp:52-231, Constraints.cpp:333-372
//   CoalesceSameConstraints        Constraints.cpp:74-150
//   ComputeNormAbund               Constraints.cpp:152-238
#pragma once
#include "psc_internal.h"
#include "std_sort_order.h"

namespace pscdev {

// --------------------------------------------------------------------------------------------
// K0: read range of every (sample, component).
// The reference streams each BAM forward through the components in genome order; read x is
// *examined* by the first component whose last subexon end is >= pos(x) on x's chromosome
// (Constraints.cpp:272-283).  Both reads and components are sorted, so the reads examined by
// component c are exactly those of its chromosome with pos in (maxEnd(c-1), maxEnd(c)].
// --------------------------------------------------------------------------------------------
PSC_DEV int upper_bound_pos(const int *pos, long long lo, long long hi, int key)
{ // first index in [lo, hi) with pos > key
	while (lo < hi) {
		long long mid = (lo + hi) >> 1;
		if (pos[mid] <= key) lo = mid + 1;
		else hi = mid;
	}
	return (int)lo;
}

PSC_DEV void group_range_one(const GraphDev &g, const SampleDev &s, int c)
{
	int chr = g.comp_chr[c];
	long long a = 0, b = 0;
	if (chr >= 0 && chr < s.n_chr) {
		a = s.chr_off[chr];
		b = s.chr_off[chr + 1];
	}
	int lo = (int)a, hi = (int)a;
	if (a < b) {
		int pe = g.comp_prev_end[c];
		lo = pe == INT32_MIN ? (int)a : upper_bound_pos(s.pos, a, b, pe);
		hi = upper_bound_pos(s.pos, a, b, g.comp_max_end[c]);
	}
	s.grp_lo[c] = lo;
	s.grp_hi[c] = hi;
	int n = g.comp_off[c + 1] - g.comp_off[c];
	// bits_off is turned into an exclusive prefix sum by the caller
	s.bits_off[c] = (long long)(hi - lo) * ((n + 63) >> 6);
}

// --------------------------------------------------------------------------------------------
// per-read conversion (Constraints.cpp:4-72 + 285-326).  Returns first (>=0) or -1.
// --------------------------------------------------------------------------------------------
PSC_DEV int convert_read(const GraphDev &g, const SampleDev &s, int c, int x, bool primaryAsUnique,
                         uint64_t *bits, int W, int4 &meta)
{
	int base = g.comp_off[c], n = g.comp_off[c + 1] - base;
	const int *sst = g.se_start + base, *sen = g.se_end + base, *pmx = g.se_pmax_end + base;
	const int8_t *lt = g.lt + base, *rt = g.rt + base;
	uint32_t so = s.seg_off[x];
	int segCnt = (int)(s.seg_off[x + 1] - so);
	const int *sa = s.seg_a + so, *sb = s.seg_b + so;
	int pos = sa[0];
	// tag: first subexon with end >= pos (Constraints.cpp:277)
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (pmx[mid] < pos) lo = mid + 1;
		else hi = mid;
	}
	int tag = lo;
	meta = make_int4(-1, -1, 0, 0);
	if (tag >= n) return -1;
	if (sb[segCnt - 1] < sst[tag]) return -1;
	uint8_t fl = s.flags[x];
	if (!(fl & 1)) return -1; // only primary alignments become constraints (Constraints.cpp:302)
	for (int w = 0; w < W; ++w) bits[w] = 0;
	int k = tag, first = -1, last = -1, cover = 0;
	for (int i = 0; i < segCnt; ++i) {
		int a = sa[i], b = sb[i];
		cover += b - a + 1;
		int leftIdx = -1, rightIdx = -1;
		for (; k < n; ++k) {
			int ss = sst[k], ee = sen[k];
			if (ss > b) break;
			if (a > ee) continue;
			bool ok = (ss >= a && ee <= b) || (i == 0 && ss < a && ee <= b) ||
			          (i == segCnt - 1 && ss >= a && ee > b) || (i == 0 && i == segCnt - 1 && ss < a && ee > b);
			if (!ok) return -1;
			if (leftIdx == -1) leftIdx = k;
			rightIdx = k;
			bt_set(bits, k);
		}
		if (leftIdx == -1) return -1;
		if (!((lt[leftIdx] == 0 || sst[leftIdx] <= a) && (rt[rightIdx] == 0 || sen[rightIdx] >= b))) return -1;
		if (i > 0) { // the intron must be an edge of the graph
			int e0 = g.next_off[base + last], e1 = g.next_off[base + last + 1];
			bool found = false;
			for (int e = e0; e < e1; ++e)
				if (g.next_idx[e] == leftIdx) { found = true; break; }
			if (!found) return -1;
		}
		for (int j = leftIdx + 1; j <= rightIdx; ++j)
			if (sst[j] > sen[j - 1] + 1) return -1;
		if (i == 0) first = leftIdx;
		last = rightIdx;
	}
	// clipped ends are only kept at true transcript ends (Constraints.cpp:310-326)
	if (fl & 2) {
		if ((first < n - 1 && sen[first] + 1 == sst[first + 1]) ||
		    g.prev_off[base + first + 1] > g.prev_off[base + first] || (sb[0] - sa[0] + 1) <= cover / 3.0)
			return -1;
	}
	if (fl & 4) {
		int t = segCnt - 1;
		if ((last > 0 && sst[last] - 1 == sen[last - 1]) || g.next_off[base + last + 1] > g.next_off[base + last] ||
		    (sb[t] - sa[t] + 1) <= cover / 3.0)
			return -1;
	}
	int uniq = primaryAsUnique ? 1 : ((fl & 16) ? 1 : 0);
	int kind = 0; // 1 left mate, 2 right mate, 3 same-position first mate
	if (fl & 32) {
		int mp = s.mate_pos[x];
		if (mp > pos) kind = 1;
		else if (mp < pos) kind = 2;
		else if (fl & 8) kind = 3;
	}
	meta = make_int4(first, last, cover, (int)s.nh[x] | (uniq << 8) | (kind << 16));
	return first;
}

PSC_DEV uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

// --------------------------------------------------------------------------------------------
// K1+K2 fused: one cooperative group builds the whole table of group (sample, comp).
//
// Equal bit vectors are found by HASHING (open addressing over the candidates' bit rows), so only the distinct
// constraints -- typically a tenth of the reads -- are sorted; the mate join is a hash join on the read-name key and
// the mate-pair list is deduplicated the same way.  Tables and the lists that get sorted live in the group's slice of
// shared memory when they fit (PSC_CG_H slots / PSC_CG_D entries), else in the group's own per-read scratch rows.
// --------------------------------------------------------------------------------------------
#define PSC_CG_H 512
#define PSC_CG_D 256
#define PSC_CG_SMEM_INTS (PSC_CG_H + PSC_CG_D)

PSC_DEV uint32_t hash_words(const uint64_t *b, int W)
{
	uint64_t h = 0x9E3779B97F4A7C15ull;
	for (int w = 0; w < W; ++w) {
		h ^= b[w] + (uint64_t)w;
		h *= 0xff51afd7ed558ccdull;
		h ^= h >> 29;
	}
	return (uint32_t)(h ^ (h >> 32));
}

// number of slots for up to `items` entries in a group of R reads: a power of two > items, never more than 2R
PSC_DEV int table_slots(int items)
{
	int H = 16;
	while (H < 2 * items) H <<= 1;
	return H;
}

// A group with NH = 3 reads: the weight sums are inexact, so they are accumulated in the order the reference's
// std::sort leaves equal constraints in, including its capacity-triggered partial coalescing
// (Constraints.cpp:378-390; the vector starts empty per component, :247-254, so capacities are 1, 2, 4, ...).
// comp(a, b) of CompSortConstraints is cid[a] < cid[b] once the rows are numbered in sorted order.
PSC_DEV void nh3_weights(const Coop &co, const SampleDev &s, int lo, int R)
{
	if (co.lane == 0) {
		const int4 *meta = s.cand_meta + lo;
		int *P = s.sortbuf + 4 * (long long)lo, *tk = P + R, *ek = s.pair_uniq_in + lo;
		double *ew = s.con_weight + lo, *tw = (double *)(s.pair_key + lo);
		int size = 0, cap = 0, thr = 16384;
		auto coalesce = [&]() {
			if (size <= 0) return;
			for (int i = 0; i < size; ++i) P[i] = i;
			std_sort_order(P, size, [&](int x, int y) { return ek[x] < ek[y]; });
			int k = 0;
			tk[0] = ek[P[0]];
			tw[0] = ew[P[0]];
			for (int i = 1; i < size; ++i) {
				int x = P[i];
				if (ek[x] == tk[k]) tw[k] += ew[x];
				else {
					++k;
					tk[k] = ek[x];
					tw[k] = ew[x];
				}
			}
			size = k + 1;
			for (int i = 0; i < size; ++i) {
				ek[i] = tk[i];
				ew[i] = tw[i];
			}
		};
		for (int r = 0; r < R; ++r) {
			int4 mt = meta[r];
			if (mt.x < 0) continue;
			if (size == cap) cap = cap ? 2 * cap : 1;
			int nh = mt.w & 0xff;
			ek[size] = s.cand_cid[lo + r];
			ew[size] = (1.0 / (double)(nh > 0 ? nh : 1)) * 10; // Constraints.cpp:294-296 with Q1
			++size;
			if (size > thr && size == cap) {
				coalesce();
				if (size >= cap / 2) thr *= 2;
			}
		}
		coalesce(); // ek[k] == k and ew[k] == con_weight[lo + k] from here on
	}
	co.sync();
}

PSC_DEV void build_group(const Coop &co, const GraphDev &g, const SampleDev &s, int c, bool primaryAsUnique, int *smem)
{
	int lo = s.grp_lo[c], hi = s.grp_hi[c];
	int R = hi - lo;
	int base = g.comp_off[c], n = g.comp_off[c + 1] - base;
	int W = (n + 63) >> 6;
	if (R <= 0) {
		if (co.lane == 0) { s.n_con[c] = 0; s.n_pair[c] = 0; }
		return;
	}
	uint64_t *cbits = s.cand_bits + s.bits_off[c];
	uint64_t *obits = s.con_bits + s.bits_off[c];
	// global scratch rows of the group: gA 4R ints (hash tables), gB 2R ints (the pair-key row, free until the mate
	// join), gC R ints
	int *gA = s.sortbuf + 4 * (long long)lo;
	int *gB = (int *)(s.pair_key + lo);
	int *gC = s.pair_uniq_in + lo;
	int *smTable = smem, *smList = smem ? smem + PSC_CG_H : nullptr;
	const int4 *meta = s.cand_meta + lo;

	// ---- A: convert every read ----
	int V = 0;
	for (int b0 = 0; b0 < R; b0 += co.n) {
		int r = b0 + co.lane;
		bool ok = false;
		if (r < R) {
			int4 mt;
			ok = convert_read(g, s, c, lo + r, primaryAsUnique, cbits + (long long)r * W, W, mt) >= 0;
			s.cand_meta[lo + r] = mt;
			s.cand_cid[lo + r] = -1;
		}
		V += popc32(co.ballot(ok));
	}
	co.sync();
	if (V == 0) {
		if (co.lane == 0) { s.n_con[c] = 0; s.n_pair[c] = 0; }
		return;
	}
	// ---- B: equal vectors share a slot of a hash table over the bit rows; the slot keeps one representative read ----
	int H = table_slots(V); // < 4V <= 4R
	int *table = (smTable && H <= PSC_CG_H) ? smTable : gA;
	for (int k = co.lane; k < H; k += co.n) table[k] = -1;
	co.sync();
	for (int b0 = 0; b0 < R; b0 += co.n) {
		int r = b0 + co.lane;
		if (r < R && meta[r].x >= 0) {
			const uint64_t *mb = cbits + (long long)r * W;
			uint32_t h = hash_words(mb, W) & (uint32_t)(H - 1);
			for (;;) {
				int old = atomic_cas(&table[h], -1, r);
				if (old == -1) break; // r represents this vector
				const uint64_t *ob = cbits + (long long)old * W;
				bool same = true;
				for (int w = 0; w < W; ++w) same = same && ob[w] == mb[w];
				if (same) break;
				h = (h + 1) & (uint32_t)(H - 1); // the table is at most half full
			}
			s.cand_cid[lo + r] = (int)h;
		}
	}
	co.sync();
	// ---- C: the distinct vectors, sorted == CompSortConstraints, are the constraint rows ----
	int C = 0;
	{
		int *dlist = gB; // room for 2R >= next_pow2(C) entries
		for (int k0 = 0; k0 < H; k0 += co.n) {
			int k = k0 + co.lane;
			bool occ = k < H && table[k] >= 0;
			uint32_t m = co.ballot(occ);
			if (occ) dlist[C + co.rank_in(m)] = k;
			C += popc32(m);
		}
		co.sync();
		if (smList && next_pow2(C) <= PSC_CG_D) {
			for (int k = co.lane; k < C; k += co.n) smList[k] = dlist[k];
			dlist = smList;
			co.sync();
		}
		coop_sort(co, dlist, C, [&](int a, int b) {
			int ra = table[a], rb = table[b];
			int fa = meta[ra].x, fb = meta[rb].x;
			if (fa != fb) return fa < fb;
			return cmp_bits_without_first(cbits + (long long)ra * W, cbits + (long long)rb * W, W) < 0;
		});
		for (int k = co.lane; k < C; k += co.n) {
			int rep = table[dlist[k]];
			for (int w = 0; w < W; ++w) obits[(long long)k * W + w] = cbits[(long long)rep * W + w];
			s.con_first[lo + k] = meta[rep].x;
			s.con_last[lo + k] = meta[rep].y;
			s.con_support[lo + k] = 0;
			s.con_uniq[lo + k] = 0;
			s.con_maxlen[lo + k] = 0;
			s.con_nhcnt[lo + k] = make_int4(0, 0, 0, 0);
		}
		co.sync();
		for (int k = co.lane; k < C; k += co.n) table[dlist[k]] = k; // slot -> row
		co.sync();
	}
	bool has3 = false;
	for (int r = co.lane; r < R; r += co.n) {
		int4 mt = meta[r];
		if (mt.x < 0) continue;
		int cid = table[s.cand_cid[lo + r]];
		s.cand_cid[lo + r] = cid;
		atomic_add(&s.con_support[lo + cid], 1);
		atomic_add(&s.con_uniq[lo + cid], (mt.w >> 8) & 1);
		atomic_max(&s.con_maxlen[lo + cid], mt.z);
		int nh = mt.w & 0xff;
		int *cnt = (int *)&s.con_nhcnt[lo + cid];
		if (nh >= 1 && nh <= 4) atomic_add(&cnt[nh - 1], 1);
		has3 = has3 || nh == 3;
	}
	has3 = co.any(has3);
	co.sync();
	if (has3) nh3_weights(co, s, lo, R);
	// ---- D: mate pairs ----
	int M = 0;
	if (s.frag_stdev == 0) {
		// single-end library: one (i, i) pair per constraint (Constraints.cpp:405-422)
		for (int k = co.lane; k < C; k += co.n) {
			s.pair_i[lo + k] = k;
			s.pair_j[lo + k] = k;
			s.pair_support[lo + k] = s.con_support[lo + k];
			s.pair_uniq[lo + k] = s.con_uniq[lo + k];
			s.pair_ec[lo + k] = 1;
		}
		M = C;
	} else {
		// MateReadIds keys an entry by (read id, matePos); the right mate asks with its own pos.  Hash join: the left
		// mates go into a table keyed by name_key ^ mix(matePos), the latest read winning a key (the reference's heap
		// hands out the last inserted entry); every right mate looks its key up.
		uint64_t *jkey = s.pair_key + lo;
		int J = 0;
		for (int b0 = 0; b0 < R; b0 += co.n) {
			int r = b0 + co.lane;
			int kind = 0;
			if (r < R && meta[r].x >= 0) kind = (meta[r].w >> 16) & 3;
			if (kind == 1 || kind == 2) {
				int jp = kind == 1 ? s.mate_pos[lo + r] : s.pos[lo + r];
				jkey[r] = s.name_key[lo + r] ^ mix64((uint64_t)(uint32_t)jp + 0x9e3779b97f4a7c15ull);
			}
			J += popc32(co.ballot(kind == 1));
		}
		H = table_slots(J);
		table = (smTable && H <= PSC_CG_H) ? smTable : gA;
		for (int k = co.lane; k < H; k += co.n) table[k] = -1;
		co.sync();
		for (int r = co.lane; r < R; r += co.n) {
			if (meta[r].x < 0 || ((meta[r].w >> 16) & 3) != 1) continue;
			const uint64_t key = jkey[r];
			uint32_t h = (uint32_t)mix64(key) & (uint32_t)(H - 1);
			for (;;) {
				int old = atomic_cas(&table[h], -1, r);
				if (old == -1) break;
				if (jkey[old] == key) { // same key: the later read wins (whoever sits in the slot has this key)
					atomic_max(&table[h], r);
					break;
				}
				h = (h + 1) & (uint32_t)(H - 1);
			}
		}
		co.sync();
		// right mates: matched left read, or -1
		for (int r = co.lane; r < R; r += co.n) {
			int left = -1;
			if (meta[r].x >= 0 && ((meta[r].w >> 16) & 3) == 2) {
				const uint64_t key = jkey[r];
				uint32_t h = (uint32_t)mix64(key) & (uint32_t)(H - 1);
				for (;;) {
					int cur = table[h];
					if (cur < 0) break;
					if (jkey[cur] == key) {
						left = cur;
						break;
					}
					h = (h + 1) & (uint32_t)(H - 1);
				}
			}
			gC[r] = left;
		}
		co.sync();
		// pair list: (cid[left], cid[me]) for matched right mates, (cid, cid) for same-position first mates.  The pair
		// keys replace the join keys (all look-ups are done), then equal pairs share a slot like equal vectors did.
		int Pn = 0;
		for (int b0 = 0; b0 < R; b0 += co.n) {
			int r = b0 + co.lane;
			bool in = false;
			uint64_t key = 0;
			if (r < R && meta[r].x >= 0) {
				int kind = (meta[r].w >> 16) & 3;
				int cj = s.cand_cid[lo + r];
				if (kind == 2) {
					int left = gC[r];
					if (left >= 0) {
						in = true;
						key = ((uint64_t)(uint32_t)s.cand_cid[lo + left] << 32) | (uint32_t)cj;
					}
				} else if (kind == 3) {
					in = true;
					key = ((uint64_t)(uint32_t)cj << 32) | (uint32_t)cj;
				}
			}
			if (r < R) {
				jkey[r] = key;
				gC[r] = in ? 1 : 0;
			}
			Pn += popc32(co.ballot(in));
		}
		co.sync();
		if (Pn > 0) {
			const uint64_t *pkey = jkey;
			H = table_slots(Pn); // < 4 Pn <= 4R
			table = (smTable && H <= PSC_CG_H) ? smTable : gA;
			for (int k = co.lane; k < H; k += co.n) table[k] = -1;
			co.sync();
			for (int r = co.lane; r < R; r += co.n) {
				if (!gC[r]) continue;
				const uint64_t key = pkey[r];
				uint32_t h = (uint32_t)mix64(key) & (uint32_t)(H - 1);
				for (;;) {
					int old = atomic_cas(&table[h], -1, r);
					if (old == -1 || pkey[old] == key) break;
					h = (h + 1) & (uint32_t)(H - 1);
				}
				gC[r] = (int)h + 1; // slot of the pair read
			}
			co.sync();
			int *dlist = (int *)(s.con_norm + lo); // 2R ints, written by step E only
			for (int k0 = 0; k0 < H; k0 += co.n) {
				int k = k0 + co.lane;
				bool occ = k < H && table[k] >= 0;
				uint32_t m = co.ballot(occ);
				if (occ) dlist[M + co.rank_in(m)] = k;
				M += popc32(m);
			}
			co.sync();
			if (smList && next_pow2(M) <= PSC_CG_D) {
				for (int k = co.lane; k < M; k += co.n) smList[k] = dlist[k];
				dlist = smList;
				co.sync();
			}
			coop_sort(co, dlist, M, [&](int a, int b) { return pkey[table[a]] < pkey[table[b]]; });
			for (int k = co.lane; k < M; k += co.n) {
				uint64_t key = pkey[table[dlist[k]]];
				s.pair_i[lo + k] = (int)(key >> 32);
				s.pair_j[lo + k] = (int)(key & 0xffffffffu);
				s.pair_ec[lo + k] = 2;
				s.pair_support[lo + k] = 0;
				s.pair_uniq[lo + k] = 0;
			}
			co.sync();
			for (int k = co.lane; k < M; k += co.n) table[dlist[k]] = k; // slot -> pair row
			co.sync();
			for (int r = co.lane; r < R; r += co.n) {
				if (!gC[r]) continue;
				int pid = table[gC[r] - 1];
				atomic_add(&s.pair_support[lo + pid], 1);
				atomic_add(&s.pair_uniq[lo + pid], (meta[r].w >> 8) & 1);
			}
		}
	}
	co.sync();
	// ---- E: weights, effective lengths, normalised abundance (Constraints.cpp:152-238) ----
	const int *sst = g.se_start + base, *sen = g.se_end + base;
	const int8_t *lt = g.lt + base, *rt = g.rt + base;
	for (int k = co.lane; k < C; k += co.n) {
		int4 nhc = s.con_nhcnt[lo + k];
		// sum of (1.0/NH)*10 over the merged reads; exact (order-free) for NH in {1,2,4}, see C2 for NH = 3
		double weight = (nhc.x * 10.0 + nhc.y * 5.0) + nhc.w * 2.5;
		if (has3) weight = s.con_weight[lo + k];
		int first = s.con_first[lo + k], last = s.con_last[lo + k];
		const uint64_t *b = obits + (long long)k * W;
		int readLen = s.con_maxlen[lo + k];
		if (first + 1 < last)
			for (int q = first + 1; q < last; ++q)
				if (bt_test(b, q)) readLen -= sen[q] - sst[q] + 1;
		int eff;
		if (first == last) {
			eff = (sen[first] - readLen + 1) - sst[first] + 1;
			if (eff <= 0) eff = (sen[first] - sst[first] + 1) / 2 + 1;
		} else {
			int start = sen[first] + 1 - (readLen - 1), end;
			if (start < sst[first]) start = sst[first];
			if (sen[last] - sst[last] + 1 >= readLen - 1 || rt[last] == 0) end = sen[first];
			else end = sen[first] + 1 - (readLen - (sen[last] - sst[last] + 1));
			if (end < start) end = sst[first];
			eff = end - start + 1;
		}
		double na = weight / (double)eff;
		if ((lt[first] == 0 && sen[first] - sst[first] + 1 >= 8 * s.read_len) ||
		    (rt[last] == 0 && sen[last] - sst[last] + 1 >= 8 * s.read_len))
			na *= 2;
		s.con_weight[lo + k] = weight;
		s.con_norm[lo + k] = na;
	}
	co.sync();
	for (int k = co.lane; k < M; k += co.n) {
		int i = s.pair_i[lo + k], j = s.pair_j[lo + k];
		double a = s.con_norm[lo + i], b = s.con_norm[lo + j];
		double v = a < b ? a : b;
		if (i != j) {
			if (lt[s.con_first[lo + i]] == 0 && s.con_first[lo + i] == s.con_last[lo + i] && a < b) v = b;
			else if (rt[s.con_last[lo + j]] == 0 && s.con_first[lo + j] == s.con_last[lo + j] && a > b) v = a;
		}
		s.pair_norm[lo + k] = v;
	}
	if (co.lane == 0) {
		s.n_con[c] = C;
		s.n_pair[c] = M;
	}
}

} // namespace pscdev
