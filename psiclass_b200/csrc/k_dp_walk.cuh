// k_dp_walk.cuh -- SolveSubTranscript / SearchSubTranscript (TranscriptDecider.cpp:374-857) as one
// iterative walk.  Included by k_dp.cuh.
//
// The reference recurses: SolveSubTranscript(key) tries every successor of the key's last subexon
// with SearchSubTranscript, which extends the walk subexon by subexon and, wherever no pending
// constraint reaches further, calls SolveSubTranscript on the walk as a new key.  Here
//   * an "owner" is one active SolveSubTranscript call (its key lives in the path array P, its
//     best-so-far path `pdp` in the fd* arrays, one slot per nesting level);
//   * the walk below an owner is a plain depth-first iteration over path positions: per position
//     only (strand, successor iterator, #pending extensions, #uncovered edges, path bits) are kept,
//     and they are written to memory only when the position is descended from or a nested owner is
//     pushed -- a leaf (memo hit or dead end), which is the common case, costs no frame traffic;
//   * path bit vectors are built incrementally (parent bits | new subexon) instead of from the path.
// Depth-first order, strand rules and the strict '>' fold are the reference's, so results are
// bit-identical (tests/: oracle parity on enumerate + DP sets).
#pragma once

namespace pscdev {

// per position: 0 strand, 1 successor iterator, 2 extCnt, 3 walkUncov
#define PPOS(p, k) pp[(p) * 4 + (k)]
// per owner: 0 vbase 1 vcnt 2 strand 3 tcStart 4 extBase 5 extCnt 6 successor iterator of the key's last node 7 keyUncov
#define OWN(o, k) of[(o) * 8 + (k)]

PSC_DEVN long long dp_solve_root(DpState &stRef, int start)
{
	// private copy: see the note in dp_pick about escaped addresses
	DpState st = stRef;
	const Coop co = st.co;
	const CompView v = st.v;
	const DpScratch d = st.d;
	const int W = st.W;
	int *const P = d.P;
	int *const pp = d.fr;
	int *const of = d.fr + (v.n + 4) * 4;
	uint64_t *const pbits = d.pbits; // [(n+4)*W] owner key | walk up to the position
	uint64_t *const kbits = d.fdkey; // [(n+3)*W] key bits per owner
	uint64_t subl[2], cbl[2];
	uint64_t *sub = W <= 2 ? subl : d.work + 2 * W; // scratch path bits (the reference's bufferTxpt)
	uint64_t *cb = W <= 2 ? cbl : d.work + 4 * W;    // bits of the child position being visited
	int extTop = 0;
	int ok = -1; // top of the owner stack
	long long ret = -1;
	++st.nRoots;

	// ---- registers describing the current owner and the current position ----
	int o_vbase = 0, o_vcnt = 0, o_strand = 0, o_tcStart = 0, o_extBase = 0, o_extCnt = 0, o_keyUncov = 0;
	long long o_mi = -1;
	int cp = 0, c_strand = 0, c_ei = 0, c_ext = 0, c_uncov = 0;
	// pending child (used across a nested solve)
	int np = 0, n_strand = 0, n_ext = 0, n_uncov = 0;
	bool keepSearch = false;

	co.sync();
	if (co.lane == 0) P[0] = start;
	co.sync();
	// request: new owner with key P[nvbase .. nvbase+nvcnt)
	int nvbase = 0, nvcnt = 1, nstrand = 0, ntcStart = 0;
	int mode = 0; // 0 enter owner, 1 next child of (owner, cp), 2 evaluate child np

	for (;;) {
		if (mode == 0) {
			// =========== SolveSubTranscript(key) ===========
			const int *key = P + nvbase;
			long long mi = dp_memo_index(st, key, nvcnt);
			DpMemo *me = d.ent + mi;
			uint64_t *mbits = d.ebits + mi * W;
			// key bits: needed for the hash probe and as the start of every path below this owner
			dp_build_sub(st, sub, nullptr, key, nvcnt, nullptr, 0);
			bool hit = false;
			if (dp_memo_hit(st, me, nstrand)) {
				if (nvcnt <= 2) hit = true;
				else if (me->cnt == nvcnt && me->first == key[0]) {
					bool same = true;
					int lo = key[0], hi = key[nvcnt - 1];
					for (int w = lo >> 6; w <= (hi >> 6) && same; ++w)
						if ((mbits[w] & keep_mask(w, lo, hi)) != sub[w]) same = false;
					hit = same;
				}
			}
			if (hit) {
				ret = mi;
				if (ok < 0) break; // the root itself was memoised
				mode = 3;          // return into the owner below
				continue;
			}
			++st.nStates;
			// ---- fresh evaluation: this key becomes the new top owner ----
			if (ok >= 0) {
				// park the owner below: its position registers go to memory
				co.sync();
				if (co.lane == 0) {
					if (cp == o_vbase + o_vcnt - 1) OWN(ok, 6) = c_ei;
					else PPOS(cp, 1) = c_ei;
					PPOS(np, 0) = n_strand;
					PPOS(np, 2) = n_ext;
					PPOS(np, 3) = n_uncov;
				}
			}
			++ok;
			o_vbase = nvbase;
			o_vcnt = nvcnt;
			o_strand = nstrand;
			o_mi = mi;
			int last = key[nvcnt - 1];
			{ // first constraint with first >= key[0]
				int lo = key[0] > 0 ? d.tcEnd[key[0] - 1] : 0;
				o_tcStart = lo > ntcStart ? lo : ntcStart;
			}
			o_keyUncov = 0;
			for (int i = 0; i + 1 < nvcnt; ++i)
				if (dp_edge_uncovered(st, key[i], key[i + 1])) ++o_keyUncov;
			double cover = -1, beta0 = 1e300;
			if (st.ce[last]) {
				bool bm = false;
				cover = dp_scan_cover(st, sub, key[0], last, o_tcStart, last, 0, bm, beta0);
			}
			// extensions: constraints whose head matches the tail of the partial path
			o_extBase = extTop;
			o_extCnt = 0;
			{
				int end = dp_tc_end(st, o_tcStart, last);
				for (int b0 = o_tcStart; b0 < end; b0 += co.n) {
					int i = b0 + co.lane;
					bool is = i < end && compat(sub, key[0], last, true, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 2;
					uint32_t m = co.ballot(is);
					if (is) {
						int p = o_extBase + o_extCnt + co.rank_in(m);
						if (p < d.extCap) d.ext[p] = i;
					}
					o_extCnt += popc32(m);
				}
				if (o_extBase + o_extCnt > d.extCap) {
					*st.err = 21;
					o_extCnt = d.extCap - o_extBase;
				}
				co.sync();
				dp_sort_ext(st, d.ext + o_extBase, o_extCnt, o_extBase + 2 * o_extCnt <= d.extCap);
			}
			extTop = o_extBase + o_extCnt;
			if (co.lane == 0) {
				uint64_t *vb = d.fdbits + (long long)ok * W, *kb = kbits + (long long)ok * W;
				for (int w = 0; w < W; ++w) { vb[w] = sub[w]; kb[w] = sub[w]; }
				d.fdcover[ok] = cover;
				d.fdbeta[ok] = beta0;
				d.fdmeta[ok * 4 + 0] = key[0];
				d.fdmeta[ok * 4 + 1] = last;
				d.fdmeta[ok * 4 + 2] = o_strand;
				OWN(ok, 0) = o_vbase; OWN(ok, 1) = o_vcnt; OWN(ok, 2) = o_strand; OWN(ok, 3) = o_tcStart;
				OWN(ok, 4) = o_extBase; OWN(ok, 5) = o_extCnt; OWN(ok, 7) = o_keyUncov;
			}
			co.sync();
			cp = o_vbase + o_vcnt - 1;
			c_strand = o_strand;
			c_ei = v.noff[last];
			c_ext = o_extCnt;
			c_uncov = o_keyUncov;
			mode = 1;
			continue;
		}
		if (mode == 3) {
			// =========== a nested SolveSubTranscript returned `ret`: resume the owner below ===========
			// (ok is the owner that issued the request; its registers may be stale only if a fresh
			// evaluation happened in between, in which case they were parked above)
			mode = 2;
			continue;
		}
		if (mode == 4) {
			// =========== the top owner is finished: store it and return into the owner below ===========
			DpMemo *me = d.ent + o_mi;
			uint64_t *mbits = d.ebits + o_mi * W;
			co.sync();
			if (co.lane == 0) {
				const uint64_t *vb = d.fdbits + (long long)ok * W;
				for (int w = 0; w < W; ++w) mbits[w] = vb[w];
				me->first = d.fdmeta[ok * 4 + 0];
				me->last = d.fdmeta[ok * 4 + 1];
				me->strand = d.fdmeta[ok * 4 + 2];
				me->cover = d.fdcover[ok];
				me->beta = d.fdbeta[ok];
				me->iter = st.iter;
				me->mode = st.forAbundance ? 1 : 0;
				me->cnt = o_vcnt;
				me->minAbundance = st.minAbundance;
				me->timeStamp = st.timeStamp;
				me->epoch = st.epoch;
			}
			co.sync();
			ret = o_mi;
			extTop = o_extBase;
			int keyEnd = o_vbase + o_vcnt - 1; // == the child position np of the owner below
			--ok;
			if (ok < 0) break;
			// reload the owner below and the position it was working at
			o_vbase = OWN(ok, 0); o_vcnt = OWN(ok, 1); o_strand = OWN(ok, 2); o_tcStart = OWN(ok, 3);
			o_extBase = OWN(ok, 4); o_extCnt = OWN(ok, 5); o_keyUncov = OWN(ok, 7);
			o_mi = dp_memo_index(st, P + o_vbase, o_vcnt);
			np = keyEnd;
			cp = np - 1;
			if (cp == o_vbase + o_vcnt - 1) {
				c_strand = o_strand; c_ei = OWN(ok, 6); c_ext = o_extCnt; c_uncov = o_keyUncov;
			} else {
				c_strand = PPOS(cp, 0); c_ei = PPOS(cp, 1); c_ext = PPOS(cp, 2); c_uncov = PPOS(cp, 3);
			}
			n_strand = PPOS(np, 0); n_ext = PPOS(np, 2); n_uncov = PPOS(np, 3);
			keepSearch = false;
			// child bits were parked in pbits[np]
			if (W <= 2) {
				cb[0] = pbits[(long long)np * W];
				if (W == 2) cb[1] = pbits[(long long)np * W + 1];
			} else {
				co.sync();
				if (co.lane == 0)
					for (int w = 0; w < W; ++w) cb[w] = pbits[(long long)np * W + w];
				co.sync();
			}
			mode = 2;
			continue;
		}
		if (mode == 1) {
			// =========== next successor of position cp (SearchSubTranscript's loops, .cpp:618-637, 798-819) ===========
			int a = P[cp];
			int b = -1, ns = c_strand;
			while (c_ei < v.noff[a + 1]) {
				int x = v.nidx[c_ei++];
				if ((same_strand(v.rs[a], c_strand) && same_strand(v.ls[x], c_strand)) || v.st[x] == v.en[a] + 1) {
					b = x;
					if (v.st[x] > v.en[a] + 1) ns = v.rs[a];
					break;
				}
			}
			if (b < 0) {
				// no successor left at cp
				if (cp == o_vbase + o_vcnt - 1) {
					mode = 4; // the owner is done
					continue;
				}
				--cp;
				if (cp == o_vbase + o_vcnt - 1) {
					c_strand = o_strand; c_ei = OWN(ok, 6); c_ext = o_extCnt; c_uncov = o_keyUncov;
				} else {
					c_strand = PPOS(cp, 0); c_ei = PPOS(cp, 1); c_ext = PPOS(cp, 2); c_uncov = PPOS(cp, 3);
				}
				continue;
			}
			// ---- visit child b at position np = cp + 1 (SearchSubTranscript entry, .cpp:385-422) ----
			++st.nFrames;
			np = cp + 1;
			n_strand = ns;
			n_uncov = c_uncov + (dp_edge_uncovered(st, a, b) ? 1 : 0);
			// path bits: (key | walk up to cp) | b
			{
				const uint64_t *src = cp == o_vbase + o_vcnt - 1 ? kbits + (long long)ok * W : pbits + (long long)cp * W;
				if (W <= 2) {
					cb[0] = src[0];
					if (W == 2) cb[1] = src[1];
					cb[b >> 6] |= (uint64_t)1 << (b & 63);
				} else {
					co.sync();
					if (co.lane == 0) {
						for (int w = 0; w < W; ++w) cb[w] = src[w];
						bt_set(cb, b);
					}
					co.sync();
				}
			}
			const int *parents = P + o_vbase;
			// drop the pending extensions that no longer fit (.cpp:407-412)
			{
				const int *ext = d.ext + o_extBase;
				int i;
				for (i = c_ext - 1; i >= 0; --i) {
					int x = ext[i];
					if (d.tcl[x] <= b ||
					    (bt_test(d.tcb + (long long)x * W, b) &&
					     compat(cb, parents[0], b, true, d.tcb + (long long)x * W, d.tcf[x], d.tcl[x], W) != 0))
						break;
				}
				n_ext = i + 1;
			}
			bool solveHere = v.noff[b + 1] > v.noff[b] && (n_ext == 0 || b >= d.tcl[d.ext[o_extBase + n_ext - 1]]);
			keepSearch = !solveHere;
			ret = -1;
			// the path array and the parked bits must be in place before anything deeper reads them
			co.sync();
			if (co.lane == 0) {
				P[np] = b;
				for (int w = 0; w < W; ++w) pbits[(long long)np * W + w] = cb[w];
			}
			co.sync();
			if (solveHere) {
				nvbase = o_vbase + o_vcnt;
				nvcnt = np - nvbase + 1;
				nstrand = ns;
				ntcStart = o_tcStart;
				mode = 0; // answered by the memo (-> mode 3 -> 2) or evaluated as a new owner (-> mode 4 -> 2)
				continue;
			}
			mode = 2;
			continue;
		}
		// =========== mode 2: evaluate child np given the sub-solution `ret` (.cpp:425-614) ===========
		{
			const int *parents = P + o_vbase;
			int pcnt = o_vcnt;
			int tag = P[np];
			uint64_t *pb = d.fdbits + (long long)ok * W; // pdp.seVector
			double vcover = -1, vbeta = 1e300;
			const uint64_t *vbits = nullptr;
			int vlast = -1;
			if (!keepSearch && ret >= 0) {
				vcover = d.ent[ret].cover;
				vbeta = d.ent[ret].beta;
				vbits = d.ebits + ret * W;
				vlast = d.ent[ret].last;
			}
			double pcover = d.fdcover[ok];
			bool belowMin = false;
			if (vcover >= 0) {
				double cover = vcover;
				// parents | sub-solution
				{
					const uint64_t *kb = kbits + (long long)ok * W;
					if (W <= 2) {
						sub[0] = kb[0] | vbits[0];
						if (W == 2) sub[1] = kb[1] | vbits[1];
					} else {
						co.sync();
						if (co.lane == 0)
							for (int w = 0; w < W; ++w) sub[w] = kb[w] | vbits[w];
						co.sync();
					}
				}
				if (!st.forAbundance && st.minAbundance > 0 && n_uncov > 0) belowMin = true;
				bool bm2 = false;
				double cbeta = vbeta;
				cover = dp_scan_cover(st, sub, parents[0], vlast, o_tcStart, parents[pcnt - 1], cover, bm2, cbeta);
				if (bm2) belowMin = true;
				bool takeBelow = belowMin && pcover == -1;
				if (takeBelow || (!(belowMin && pcover == -1) && cover > pcover)) {
					co.sync();
					if (co.lane == 0) {
						d.fdcover[ok] = takeBelow ? -2 : cover;
						d.fdbeta[ok] = cbeta;
						for (int w = 0; w < W; ++w) pb[w] = sub[w];
						d.fdmeta[ok * 4 + 0] = parents[0];
						d.fdmeta[ok * 4 + 1] = vlast;
						d.fdmeta[ok * 4 + 2] = n_strand;
					}
					co.sync();
					pcover = takeBelow ? -2 : cover;
				}
			} else if (vcover == -2 && pcover == -1) {
				co.sync();
				if (co.lane == 0) {
					const uint64_t *kb = kbits + (long long)ok * W;
					for (int w = 0; w < W; ++w) pb[w] = vbits[w] | kb[w];
					d.fdcover[ok] = -2;
					d.fdmeta[ok * 4 + 0] = parents[0];
					d.fdmeta[ok * 4 + 1] = vlast;
					d.fdmeta[ok * 4 + 2] = n_strand;
				}
				co.sync();
				pcover = -2;
			}
			// the path may also stop at this subexon (.cpp:524-614): path bits == cb
			if (st.ce[tag] && (vcover < 0 || st.forAbundance)) {
				bool below = false;
				double cover = 0;
				if (st.forAbundance || st.minAbundance > 0) {
					if (n_uncov > 0) below = true;
					if (st.forAbundance && below) cover = 1e-6;
				}
				bool bm2 = false;
				double cbeta = 1e300;
				cover = dp_scan_cover(st, cb, parents[0], tag, o_tcStart, tag, cover, bm2, cbeta);
				if (bm2) below = true;
				bool takeBelow = below && pcover == -1;
				if (takeBelow || (!(below && pcover == -1) && cover > pcover)) {
					co.sync();
					if (co.lane == 0) {
						d.fdcover[ok] = takeBelow ? -2 : cover;
						d.fdbeta[ok] = cbeta;
						for (int w = 0; w < W; ++w) pb[w] = cb[w];
						d.fdmeta[ok * 4 + 0] = parents[0];
						d.fdmeta[ok * 4 + 1] = tag;
						d.fdmeta[ok * 4 + 2] = n_strand;
					}
					co.sync();
				}
			}
			if (keepSearch) {
				// descend: park cp's iterator, make np the current position
				co.sync();
				if (co.lane == 0) {
					if (cp == o_vbase + o_vcnt - 1) OWN(ok, 6) = c_ei;
					else PPOS(cp, 1) = c_ei;
					PPOS(np, 0) = n_strand;
					PPOS(np, 2) = n_ext;
					PPOS(np, 3) = n_uncov;
				}
				co.sync();
				cp = np;
				c_strand = n_strand;
				c_ei = v.noff[tag];
				c_ext = n_ext;
				c_uncov = n_uncov;
			}
			ret = -1;
			mode = 1;
		}
	}
	co.sync();
	stRef.nStates = st.nStates;
	stRef.nFrames = st.nFrames;
	stRef.nRoots = st.nRoots;
	return ret;
}
#undef PPOS
#undef OWN

} // namespace pscdev
