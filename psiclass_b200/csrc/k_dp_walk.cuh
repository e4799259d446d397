// k_dp_walk.cuh -- SolveSubTranscript / SearchSubTranscript (TranscriptDecider.cpp:374-857) as one
// iterative walk.  Included by k_dp.cuh.
//
// The reference recurses: SolveSubTranscript(key) tries every successor of the key's last subexon
// with SearchSubTranscript, which extends the walk subexon by subexon and, wherever no pending
// constraint reaches further, calls SolveSubTranscript on the walk as a new key.  Here
//   * an "owner" is one active SolveSubTranscript call.  The top owner lives in registers (key
//     range, scan start, extension list, best-so-far path `pdp`); it is written to its slot in the
//     fd*/OWN arrays only when a nested owner is pushed on top of it;
//   * the walk below an owner is a plain depth-first iteration over path positions.  The current
//     position and the child being visited are registers too: (node, strand, successor iterator,
//     #pending extensions, #uncovered edges, running key hash, path bits).  They go to memory only when
//     the walk descends or a nested owner is pushed -- a leaf (memo hit or dead end), the common
//     case, touches no frame memory at all;
//   * everything a nested key needs is derived incrementally from the walk: its bits are
//     (child bits & ~owner key bits), its memo hash is the running hash, its count of uncovered
//     edges a difference of two running counts;
//   * components of up to 256 subexons (W <= NW = 2 or 4 words) keep all bit vectors in registers and use the
//     loop-free compat_r / two-constraints-per-lane scans; wider components (NW = 0) go through memory.
// Depth-first order, strand rules and the strict '>' fold are the reference's, so results are
// bit-identical (tests/: oracle parity on enumerate + DP sets).
#pragma once

namespace pscdev {

// per walk position: 0 strand, 1 successor iterator, 2 extCnt, 3 walkUncov, 4 running key hash
#define PPOS(p, k) pp[(p) * 5 + (k)]
// per owner: 0 vbase 1 vcnt 2 strand 3 tcStart 4 extBase 5 extCnt 6 successor iterator of the key's last node
//            7 keyUncov 8 memo index 9 strand of the best path
#define OWN(o, k) of[(o) * 10 + (k)]

template <int NW> PSC_DEVN long long dp_solve_root_t(DpState &stRef, int start)
{
	constexpr bool SMALL = NW > 0;
	typedef BitsR<NW> BV;
	// private copy: see the note in dp_pick about escaped addresses
	DpState st = stRef;
	const Coop co = st.co;
	const CompView v = st.v;
	const DpScratch d = st.d;
	const int W = st.W;
	const uint32_t nn = (uint32_t)v.n;
	int *const P = d.P;
	int *const pp = d.fr;
	int *const of = d.fr + (v.n + 4) * 5;
	uint64_t *const pbits = d.pbits; // [(n+4)*W] owner key | walk up to the position
	uint64_t *const kbits = d.fdkey; // [(n+3)*W] key bits per owner
	// wide components: scratch path bits (the reference's bufferTxpt) and the bits of the child being visited
	uint64_t *const sub = d.work + 2 * W;
	uint64_t *const cb = d.work + 4 * W;
	const int e0 = v.noff[0];
	int extTop = 0;
	int ok = -1; // top of the owner stack
	long long ret = -1;
	int nStates = 0, nFrames = 0; // PSC_DEBUG_DP counters

	// ---- the top owner ----
	int o_vbase = 0, o_vcnt = 0, o_strand = 0, o_tcStart = 0, o_extBase = 0, o_extCnt = 0, o_keyUncov = 0, o_ei = 0;
	int o_first = 0, o_last = 0;
	long long o_mi = -1;
	BV o_k;                              // key bits
	double o_cover = -1, o_beta = 1e300; // pdp
	BV o_p;
	int o_plast = 0, o_pstrand = 0;
	// ---- the current position and the child being visited ----
	int cp = 0, c_node = 0, c_strand = 0, c_ei = 0, c_ext = 0, c_uncov = 0;
	uint32_t c_hash = 0;
	BV c_b;
	int np = 0, n_node = 0, n_strand = 0, n_ext = 0, n_uncov = 0;
	uint32_t n_hash = 0;
	BV n_b;
	bool keepSearch = false;

	// request: new owner with key P[nvbase .. nvbase+nvcnt)
	int nvbase = 0, nvcnt = 1, nstrand = 0, ntcStart = 0, nfirst = start, nlast = start, nkeyUncov = 0;
	uint32_t nhash = 0;
	BV nk;
	nk.clear();
	nk.set(start);
	int mode = 0; // 0 enter owner, 1 next child of (owner, cp), 2 evaluate child np, 4 owner done

	co.sync();
	if (co.lane == 0) P[0] = start;
	co.sync();

	for (;;) {
		if (mode == 0) {
			// =========== SolveSubTranscript(key) ===========
			const int *key = P + nvbase;
			long long mi;
			if (nvcnt == 1) mi = nlast;
			else if (nvcnt == 2) mi = d.f2Base + (long long)nfirst * v.n + nlast; // seCnt <= 10000 is enforced by the host
			else mi = d.hashBase + nhash;
			DpMemo *me = d.ent + mi;
			uint64_t *mbits = d.ebits + mi * W;
			if (!SMALL) dp_build_sub(st, sub, nullptr, key, nvcnt, nullptr, 0);
			bool hit = false;
			if (dp_memo_hit(st, mi, nstrand)) {
				if (nvcnt <= 2) hit = true;
				else if (me->cnt == nvcnt && me->first == nfirst) {
					if (SMALL) {
						BV r, mb;
						r.range(nfirst, nlast);
						mb.load(mbits, W);
						hit = mb.masked_equals(r, nk);
					} else {
						bool same = true;
						for (int w = nfirst >> 6; w <= (nlast >> 6) && same; ++w)
							if ((mbits[w] & keep_mask(w, nfirst, nlast)) != sub[w]) same = false;
						hit = same;
					}
				}
			}
			if (hit) {
				ret = mi;
				if (ok < 0) break; // the root itself was memoised
				mode = 2;          // return into the owner
				continue;
			}
			++nStates;
			// ---- fresh evaluation: this key becomes the new top owner ----
			if (ok >= 0) {
				// park the owner below, its current position and the child that became the key's end
				if (cp == o_vbase + o_vcnt - 1) o_ei = c_ei;
				co.sync();
				if (co.lane == 0) {
					if (cp != o_vbase + o_vcnt - 1) PPOS(cp, 1) = c_ei;
					OWN(ok, 6) = o_ei;
					OWN(ok, 9) = o_pstrand;
					d.fdcover[ok] = o_cover;
					d.fdbeta[ok] = o_beta;
					d.fdmeta[ok * 4 + 1] = o_plast;
					if (SMALL) {
						o_p.store(d.fdbits + (long long)ok * W, W);
						n_b.store(pbits + (long long)np * W, W);
					}
					P[np] = n_node;
					PPOS(np, 0) = n_strand;
					PPOS(np, 2) = n_ext;
					PPOS(np, 3) = n_uncov;
					PPOS(np, 4) = (int)n_hash;
				}
			}
			++ok;
			o_vbase = nvbase;
			o_vcnt = nvcnt;
			o_strand = nstrand;
			o_mi = mi;
			o_first = nfirst;
			o_last = nlast;
			o_keyUncov = nkeyUncov;
			o_k = nk;
			{ // first constraint with first >= key[0]
				int lo = o_first > 0 ? d.tcEnd[o_first - 1] : 0;
				o_tcStart = lo > ntcStart ? lo : ntcStart;
			}
			double cover = -1, beta0 = 1e300;
			if (st.ce[o_last]) {
				bool bm = false;
				if (SMALL) cover = dp_scan_cover_r<NW>(st, o_k, o_first, o_last, o_tcStart, o_last, 0, bm, beta0);
				else cover = dp_scan_cover(st, sub, o_first, o_last, o_tcStart, o_last, 0, bm, beta0);
			}
			// extensions: constraints whose head matches the tail of the partial path
			o_extBase = extTop;
			o_extCnt = 0;
			{
				int end = dp_tc_end(st, o_tcStart, o_last);
				if (SMALL) {
					for (int b0 = o_tcStart; b0 < end; b0 += co.n) {
						int i0 = b0 + co.lane;
						bool v0 = i0 < end;
						int j0 = v0 ? i0 : end - 1;
						BV x0;
						x0.load(d.tcb + (long long)j0 * W, W);
						bool is0 = v0 && compat_r<NW>(o_k, o_first, o_last, true, x0, d.tcf[j0], d.tcl[j0]) == 2;
						uint32_t m0 = co.ballot(is0);
						if (is0) {
							int p = o_extBase + o_extCnt + co.rank_in(m0);
							if (p < d.extCap) d.ext[p] = i0;
						}
						o_extCnt += popc32(m0);
					}
				} else {
					for (int b0 = o_tcStart; b0 < end; b0 += co.n) {
						int i = b0 + co.lane;
						bool is = i < end && compat(sub, o_first, o_last, true, d.tcb + (long long)i * W, d.tcf[i], d.tcl[i], W) == 2;
						uint32_t m = co.ballot(is);
						if (is) {
							int p = o_extBase + o_extCnt + co.rank_in(m);
							if (p < d.extCap) d.ext[p] = i;
						}
						o_extCnt += popc32(m);
					}
				}
				if (o_extBase + o_extCnt > d.extCap) {
					*st.err = 21;
					o_extCnt = d.extCap - o_extBase;
				}
				co.sync();
				dp_sort_ext(st, d.ext + o_extBase, o_extCnt, o_extBase + 2 * o_extCnt <= d.extCap);
			}
			extTop = o_extBase + o_extCnt;
			// pdp starts as the key itself
			o_cover = cover;
			o_beta = beta0;
			o_p = o_k;
			o_plast = o_last;
			o_pstrand = o_strand;
			if (co.lane == 0) {
				if (SMALL) {
					o_k.store(kbits + (long long)ok * W, W);
				} else {
					uint64_t *vb = d.fdbits + (long long)ok * W, *kb = kbits + (long long)ok * W;
					for (int w = 0; w < W; ++w) { vb[w] = sub[w]; kb[w] = sub[w]; }
				}
				OWN(ok, 0) = o_vbase; OWN(ok, 1) = o_vcnt; OWN(ok, 2) = o_strand; OWN(ok, 3) = o_tcStart;
				OWN(ok, 4) = o_extBase; OWN(ok, 5) = o_extCnt; OWN(ok, 7) = o_keyUncov; OWN(ok, 8) = (int)o_mi;
			}
			co.sync();
			cp = o_vbase + o_vcnt - 1;
			c_node = o_last;
			c_strand = o_strand;
			c_ei = v.noff[o_last];
			c_ext = o_extCnt;
			c_uncov = o_keyUncov;
			c_hash = 0;
			c_b = o_k;
			mode = 1;
			continue;
		}
		if (mode == 4) {
			// =========== the top owner is finished: store it and return into the owner below ===========
			DpMemo *me = d.ent + o_mi;
			uint64_t *mbits = d.ebits + o_mi * W;
			co.sync();
			if (co.lane == 0) {
				if (SMALL) {
					o_p.store(mbits, W);
				} else {
					const uint64_t *vb = d.fdbits + (long long)ok * W;
					for (int w = 0; w < W; ++w) mbits[w] = vb[w];
				}
				me->first = o_first;
				me->last = o_plast;
				me->strand = o_pstrand;
				me->cover = o_cover;
				me->beta = o_beta;
				me->iter = st.iter;
				me->mode = st.forAbundance ? 1 : 0;
				me->cnt = o_vcnt;
				me->minAbundance = st.minAbundance;
				me->timeStamp = st.timeStamp;
				d.eepoch[o_mi] = st.epoch;
			}
			co.sync();
			ret = o_mi;
			extTop = o_extBase;
			int keyEnd = o_vbase + o_vcnt - 1; // == the child position np of the owner below
			--ok;
			if (ok < 0) break;
			// reload the owner below, the position it was working at and the child that was the key's end
			o_vbase = OWN(ok, 0); o_vcnt = OWN(ok, 1); o_strand = OWN(ok, 2); o_tcStart = OWN(ok, 3);
			o_extBase = OWN(ok, 4); o_extCnt = OWN(ok, 5); o_ei = OWN(ok, 6); o_keyUncov = OWN(ok, 7);
			o_mi = OWN(ok, 8); o_pstrand = OWN(ok, 9);
			o_first = P[o_vbase];
			o_last = P[o_vbase + o_vcnt - 1];
			o_cover = d.fdcover[ok];
			o_beta = d.fdbeta[ok];
			o_plast = d.fdmeta[ok * 4 + 1];
			if (SMALL) {
				o_k.load(kbits + (long long)ok * W, W);
				o_p.load(d.fdbits + (long long)ok * W, W);
			}
			np = keyEnd;
			n_node = P[np];
			n_strand = PPOS(np, 0); n_ext = PPOS(np, 2); n_uncov = PPOS(np, 3); n_hash = (uint32_t)PPOS(np, 4);
			cp = np - 1;
			if (cp == o_vbase + o_vcnt - 1) {
				c_node = o_last; c_strand = o_strand; c_ei = o_ei; c_ext = o_extCnt; c_uncov = o_keyUncov; c_hash = 0;
				c_b = o_k;
			} else {
				c_node = P[cp];
				c_strand = PPOS(cp, 0); c_ei = PPOS(cp, 1); c_ext = PPOS(cp, 2); c_uncov = PPOS(cp, 3); c_hash = (uint32_t)PPOS(cp, 4);
				if (SMALL) {
					c_b.load(pbits + (long long)cp * W, W);
				}
			}
			keepSearch = false;
			if (SMALL) {
				n_b.load(pbits + (long long)np * W, W);
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
			const int a = c_node;
			int b = -1, ns = c_strand;
			{
				const int eEnd = v.noff[a + 1], rsa = v.rs[a], ena = v.en[a];
				const bool aOk = same_strand(rsa, c_strand);
				while (c_ei < eEnd) {
					int x = v.nidx[c_ei++];
					int stx = v.st[x];
					if ((aOk && same_strand(v.ls[x], c_strand)) || stx == ena + 1) {
						b = x;
						if (stx > ena + 1) ns = rsa;
						break;
					}
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
					c_node = o_last; c_strand = o_strand; c_ei = o_ei; c_ext = o_extCnt; c_uncov = o_keyUncov; c_hash = 0;
					c_b = o_k;
				} else {
					c_node = P[cp];
					c_strand = PPOS(cp, 0); c_ei = PPOS(cp, 1); c_ext = PPOS(cp, 2); c_uncov = PPOS(cp, 3); c_hash = (uint32_t)PPOS(cp, 4);
					if (SMALL) {
						c_b.load(pbits + (long long)cp * W, W);
					}
				}
				continue;
			}
			// ---- visit child b at position np = cp + 1 (SearchSubTranscript entry, .cpp:385-422) ----
			++nFrames;
			np = cp + 1;
			n_node = b;
			n_strand = ns;
			n_uncov = c_uncov + (d.uncov[c_ei - 1 - e0] ? 1 : 0); // the edge (a, b) is the one the iterator just passed
			n_hash = (c_hash * nn + (uint32_t)b) % PSC_DP_HASH;
			// path bits: (key | walk up to cp) | b
			if (SMALL) {
				n_b = c_b;
				n_b.set(b);
			} else {
				const uint64_t *src = cp == o_vbase + o_vcnt - 1 ? kbits + (long long)ok * W : pbits + (long long)cp * W;
				co.sync();
				if (co.lane == 0) {
					for (int w = 0; w < W; ++w) cb[w] = src[w];
					bt_set(cb, b);
					// the path array and the parked bits must be in place before anything deeper reads them
					P[np] = b;
					for (int w = 0; w < W; ++w) pbits[(long long)np * W + w] = cb[w];
				}
				co.sync();
			}
			// drop the pending extensions that no longer fit (.cpp:407-412)
			int keptLast = 0; // .last of the furthest-reaching extension that stays pending
			{
				const int *ext = d.ext + o_extBase;
				int i;
				for (i = c_ext - 1; i >= 0; --i) {
					int x = ext[i];
					int xl = d.tcl[x];
					keptLast = xl;
					if (xl <= b) break;
					if (SMALL) {
						BV xv;
						xv.load(d.tcb + (long long)x * W, W);
						if (xv.test(b) && compat_r<NW>(n_b, o_first, b, true, xv, d.tcf[x], xl) != 0) break;
					} else {
						if (bt_test(d.tcb + (long long)x * W, b) &&
						    compat(cb, o_first, b, true, d.tcb + (long long)x * W, d.tcf[x], xl, W) != 0)
							break;
					}
				}
				n_ext = i + 1;
			}
			bool solveHere = v.noff[b + 1] > v.noff[b] && (n_ext == 0 || b >= keptLast);
			keepSearch = !solveHere;
			ret = -1;
			if (solveHere) {
				nvbase = o_vbase + o_vcnt;
				nvcnt = np - nvbase + 1;
				nstrand = ns;
				ntcStart = o_tcStart;
				nlast = b;
				nfirst = nvcnt == 1 ? b : P[nvbase];
				nhash = n_hash;
				nkeyUncov = nvcnt == 1 ? 0 : n_uncov - PPOS(nvbase, 3);
				nk.and_not(n_b, o_k);
				mode = 0; // answered by the memo (-> mode 2) or evaluated as a new owner (-> mode 4 -> 2)
				continue;
			}
			mode = 2;
			continue;
		}
		// =========== mode 2: evaluate child np given the sub-solution `ret` (.cpp:425-614) ===========
		{
			const int tag = n_node;
			uint64_t *pb = d.fdbits + (long long)ok * W; // pdp.seVector of wide components
			double vcover = -1, vbeta = 1e300;
			const uint64_t *vbits = nullptr;
			int vlast = -1;
			if (!keepSearch && ret >= 0) {
				vcover = d.ent[ret].cover;
				vbeta = d.ent[ret].beta;
				vbits = d.ebits + ret * W;
				vlast = d.ent[ret].last;
			}
			bool belowMin = false;
			if (vcover >= 0) {
				double cover = vcover;
				// parents | sub-solution
				BV sv;
				if (SMALL) {
					sv.load(vbits, W);
					sv.or_with(o_k);
				} else {
					const uint64_t *kb = kbits + (long long)ok * W;
					co.sync();
					if (co.lane == 0)
						for (int w = 0; w < W; ++w) sub[w] = kb[w] | vbits[w];
					co.sync();
				}
				if (!st.forAbundance && st.minAbundance > 0 && n_uncov > 0) belowMin = true;
				bool bm2 = false;
				double cbeta = vbeta;
				if (SMALL) cover = dp_scan_cover_r<NW>(st, sv, o_first, vlast, o_tcStart, o_last, cover, bm2, cbeta);
				else cover = dp_scan_cover(st, sub, o_first, vlast, o_tcStart, o_last, cover, bm2, cbeta);
				if (bm2) belowMin = true;
				bool takeBelow = belowMin && o_cover == -1;
				if (takeBelow || (!(belowMin && o_cover == -1) && cover > o_cover)) {
					o_cover = takeBelow ? -2 : cover;
					o_beta = cbeta;
					o_plast = vlast;
					o_pstrand = n_strand;
					if (SMALL) {
						o_p = sv;
					} else {
						co.sync();
						if (co.lane == 0)
							for (int w = 0; w < W; ++w) pb[w] = sub[w];
						co.sync();
					}
				}
			} else if (vcover == -2 && o_cover == -1) {
				o_cover = -2;
				o_plast = vlast;
				o_pstrand = n_strand;
				if (SMALL) {
					o_p.load(vbits, W);
					o_p.or_with(o_k);
				} else {
					co.sync();
					if (co.lane == 0) {
						const uint64_t *kb = kbits + (long long)ok * W;
						for (int w = 0; w < W; ++w) pb[w] = vbits[w] | kb[w];
					}
					co.sync();
				}
			}
			// the path may also stop at this subexon (.cpp:524-614): path bits == the child's bits
			if (st.ce[tag] && (vcover < 0 || st.forAbundance)) {
				bool below = false;
				double cover = 0;
				if (st.forAbundance || st.minAbundance > 0) {
					if (n_uncov > 0) below = true;
					if (st.forAbundance && below) cover = 1e-6;
				}
				bool bm2 = false;
				double cbeta = 1e300;
				if (SMALL) cover = dp_scan_cover_r<NW>(st, n_b, o_first, tag, o_tcStart, tag, cover, bm2, cbeta);
				else cover = dp_scan_cover(st, cb, o_first, tag, o_tcStart, tag, cover, bm2, cbeta);
				if (bm2) below = true;
				bool takeBelow = below && o_cover == -1;
				if (takeBelow || (!(below && o_cover == -1) && cover > o_cover)) {
					o_cover = takeBelow ? -2 : cover;
					o_beta = cbeta;
					o_plast = tag;
					o_pstrand = n_strand;
					if (SMALL) {
						o_p = n_b;
					} else {
						co.sync();
						if (co.lane == 0)
							for (int w = 0; w < W; ++w) pb[w] = cb[w];
						co.sync();
					}
				}
			}
			if (keepSearch) {
				// descend: park cp's iterator, make np the current position
				if (cp == o_vbase + o_vcnt - 1) o_ei = c_ei;
				co.sync();
				if (co.lane == 0) {
					if (cp != o_vbase + o_vcnt - 1) PPOS(cp, 1) = c_ei;
					P[np] = n_node;
					PPOS(np, 0) = n_strand;
					PPOS(np, 2) = n_ext;
					PPOS(np, 3) = n_uncov;
					PPOS(np, 4) = (int)n_hash;
					if (SMALL) {
						n_b.store(pbits + (long long)np * W, W);
					}
				}
				co.sync();
				cp = np;
				c_node = n_node;
				c_strand = n_strand;
				c_ei = v.noff[tag];
				c_ext = n_ext;
				c_uncov = n_uncov;
				c_hash = n_hash;
				c_b = n_b;
			}
			ret = -1;
			mode = 1;
		}
	}
	co.sync();
	stRef.nStates += nStates;
	stRef.nFrames += nFrames;
	++stRef.nRoots;
	return ret;
}
#undef PPOS
#undef OWN

PSC_DEV long long dp_solve_root(DpState &st, int start)
{
	if (st.W <= 2) return dp_solve_root_t<2>(st, start);
	if (st.W <= 4) return dp_solve_root_t<4>(st, start);
	return dp_solve_root_t<0>(st, start);
}

} // namespace pscdev
