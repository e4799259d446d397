// k_solve2.cuh -- AbundanceEstimation, RefineTranscripts, OutputTranscript and the per-unit phase
// drivers of Solve (TranscriptDecider.cpp:2121-2837, 3-111, 3267-3443).
#pragma once

namespace pscdev {

// the transcripts of one (component, sample) unit: indices into the candidate table + state
struct PredView {
	int T;
	UnitPred pr;
	const CandView *cd;
	PSC_DEV const uint64_t *bits(int k) const { return cd->bits + (long long)pr.id[k] * cd->W; }
	PSC_DEV int first(int k) const { return cd->first[pr.id[k]]; }
	PSC_DEV int last(int k) const { return cd->last[pr.id[k]]; }
};

PSC_DEV bool con_in_tx(const TableView &t, int k, const uint64_t *tb, int tf, int tl, int W)
{
	return compat(tb, tf, tl, false, t.bits + (long long)k * W, t.first[k], t.last[k], W) != 0;
}
// a two-subexon "chain" constraint {a, b} inside transcript tb? (IsConstraintInTranscript on tmpC)
PSC_DEV bool chain_in_tx(const uint64_t *tb, int tf, int tl, int a, int b, int W)
{
	// c.first = a, c.last = b, bits {a, b}; transcript not partial
	if (a < tf || a > tl || !bt_test(tb, a) || !bt_test(tb, b)) return false;
	if (b > tl) return false;
	// masked equality on [a, b]: the transcript must have exactly bits a and b in that range
	for (int q = a + 1; q < b; ++q)
		if (bt_test(tb, q)) return false;
	return true;
}

// ---------------------------------------------------------------------------------------------
// AbundanceEstimation (TranscriptDecider.cpp:2121-2251).  The T x M constraintsSupport matrix is
// never materialised: per pair only the sum of rho over compatible transcripts is needed, and
// the per-transcript sums run over pairs in index order, exactly like the reference.
// ---------------------------------------------------------------------------------------------
PSC_DEV void abundance_em(const Coop &co, const CompView &v, const TableView &t, const SolveParams &prm,
                          UnitScratch &sc, PredView &pv)
{
	int T = pv.T, M = t.M, W = v.W;
	if (T <= 0) return;
	int MW = (M + 63) >> 6;
	for (int i = 0; i < T; ++i)
		build_pair_row(co, t, pv.bits(i), pv.first(i), pv.last(i), W, M, MW, sc.ebtable + (long long)i * MW, sc.flagsA);
	co.sync();
	for (int i = co.lane; i < T; i += co.n) {
		int len = tx_len(v, pv.bits(i), pv.first(i), pv.last(i));
		int L = len - prm.fragLen + 2 * prm.fragStdev;
		if (L < 1) L = 1;
		sc.tlen[i] = L;
		double r = pv.pr.abundance[i] / L;
		if (pv.pr.corr[i] == -1 && r > 0.1 / (double)prm.readLen) r = 0.1 / (double)prm.readLen;
		sc.rho[i] = r;
	}
	co.sync();
	int iterCnt = 0;
	for (;;) {
		// E-step: sum of rho over the transcripts compatible with pair j (ascending i)
		for (int j = co.lane; j < M; j += co.n) {
			double sum = 0;
			for (int i = 0; i < T; ++i)
				if (bt_test(sc.ebtable + (long long)i * MW, j)) sum += sc.rho[i];
			sc.sumj[j] = sum;
		}
		co.sync();
		// M-step: lanes over transcripts, pairs in index order
		double diff = 0;
		for (int i = co.lane; i < T; i += co.n) {
			const uint64_t *bt = sc.ebtable + (long long)i * MW;
			double newAbund = 0, r = sc.rho[i];
			for (int w = 0; w < MW; ++w) { // compatible pairs only, still in index order
				uint64_t word = bt[w];
				while (word) {
					int j = w * 64 + ctz64(word);
					word &= word - 1;
					double factor = t.pec[j];
					newAbund += (r / sc.sumj[j]) * t.psup[j] * factor;
				}
			}
			pv.pr.abundance[i] = newAbund; // value of the final loop survives (.cpp:2227-2234)
			double nr = newAbund / sc.tlen[i];
			if (pv.pr.corr[i] == -1 && nr > 0.1 / (double)prm.readLen) nr = 0.1 / (double)prm.readLen;
			double tmp = r - nr;
			sc.rho[i] = nr;
			sc.dtmp[i] = tmp < 0 ? -tmp : tmp; // summed in transcript order below (fp order)
		}
		co.sync();
		for (int i = 0; i < T; ++i) diff += sc.dtmp[i];
		co.sync();
		if (diff < 1e-3) break;
		++iterCnt;
		if (iterCnt >= 1000) break;
	}
}

PSC_DEV int remove_negative(const Coop &co, PredView &pv)
{
	// RemoveNegativeAbundTranscripts (TranscriptDecider.hpp:437-454): stable compaction
	int j = 0;
	co.sync();
	if (co.lane == 0) {
		for (int i = 0; i < pv.T; ++i) {
			if (pv.pr.abundance[i] < 0) continue;
			if (i != j) {
				pv.pr.id[j] = pv.pr.id[i];
				pv.pr.abundance[j] = pv.pr.abundance[i];
				pv.pr.fpkm[j] = pv.pr.fpkm[i];
				pv.pr.corr[j] = pv.pr.corr[i];
			}
			++j;
		}
	}
	j = co.bcast(j, 0);
	co.sync();
	pv.T = j;
	return j;
}

// ---------------------------------------------------------------------------------------------
// RefineTranscripts (TranscriptDecider.cpp:2253-2837)
// ---------------------------------------------------------------------------------------------
PSC_DEV int refine_transcripts(const Coop &co, const CompView &v, const TableView &t, const SolveParams &prm,
                               const SolveDev &sv, UnitScratch &sc, PredView &pv, bool aggressive)
{
	int T = pv.T;
	if (T == 0) return 0;
	int W = v.W, M = t.M, C = t.C;
	int ng = v.usedGene - v.baseGene;
	const int *defGene = sv.def_gene + 2 * v.c;
	int readLen = prm.readLen;
	UnitPred &pr = pv.pr;
	// ---- per-gene maxima (.cpp:2287-2309), sequential ----
	if (co.lane == 0) {
		for (int g = 0; g < ng; ++g) { sc.gMaxFPKM[g] = 0; sc.gNonOvl[g] = 0; sc.gMaxCov[g] = 0; sc.gTag[g] = 0; sc.gRecover[g] = 0; }
		for (int i = 0; i < T; ++i) {
			int gid = gene_id_of(v, defGene, pv.bits(i), pv.first(i), pv.last(i));
			int len = len_from_fpkm(prm, pr.abundance[i], pr.fpkm[i]);
			int gi = gid - v.baseGene;
			if (pr.fpkm[i] > sc.gMaxFPKM[gi]) { sc.gMaxFPKM[gi] = pr.fpkm[i]; sc.gTag[gi] = i; }
			if (pr.abundance[i] * readLen / len > sc.gMaxCov[gi]) sc.gMaxCov[gi] = (pr.abundance[i] * readLen) / len;
			sc.tgid[i] = gid;
		}
		for (int i = 0; i < T; ++i) {
			int tag = sc.tgid[i] - v.baseGene;
			int mt = sc.gTag[tag];
			if ((pv.last(i) < pv.first(mt) || pv.first(i) > pv.last(mt)) && pr.fpkm[i] > sc.gNonOvl[tag])
				sc.gNonOvl[tag] = pr.fpkm[i];
		}
	}
	co.sync();
	// ---- transcripts that alone explain well-supported unique pairs get a x4 boost (.cpp:2313-2351) ----
	if (!aggressive) {
		for (int i = co.lane; i < T; i += co.n) sc.icnt[i] = 0;
		co.sync();
		for (int j = co.lane; j < M; j += co.n) {
			if (t.puniq[j] <= 5) continue;
			int cnt = 0, tag = 0;
			for (int i = 0; i < T; ++i) {
				const uint64_t *tb = pv.bits(i);
				if (con_in_tx(t, t.pi[j], tb, pv.first(i), pv.last(i), W) && con_in_tx(t, t.pj[j], tb, pv.first(i), pv.last(i), W)) {
					tag = i;
					++cnt;
				}
				if (cnt >= 2) break;
			}
			if (cnt == 1) atomic_add(&sc.icnt[tag], 1);
		}
		co.sync();
		for (int i = co.lane; i < T; i += co.n)
			if (sc.icnt[i] >= 2) {
				pr.abundance[i] *= 4;
				pr.fpkm[i] *= 4;
			}
		co.sync();
	}
	// ---- relative FPKM filter + non-overlap rescue (.cpp:2353-2418) ----
	if (co.lane == 0) {
		double filterFactor = 1.0;
		for (int i = 0; i < T; ++i)
			if (pr.fpkm[i] < filterFactor * prm.fpkmFraction * sc.gMaxFPKM[sc.tgid[i] - v.baseGene])
				pr.abundance[i] = -pr.abundance[i];
		int nres = 0;
		for (int i = 0; i < T; ++i) {
			if (pr.abundance[i] >= 0) continue;
			int j;
			for (j = 0; j < T; ++j) {
				if (pr.abundance[j] < 0 || sc.tgid[i] != sc.tgid[j]) continue;
				if (pv.first(i) <= pv.last(j) && pv.last(i) >= pv.first(j)) break;
			}
			if (j >= T && pr.fpkm[i] >= prm.fpkmFraction * sc.gNonOvl[sc.tgid[i] - v.baseGene]) sc.recList[nres++] = i;
		}
		for (int q = 0; q < nres; ++q) pr.abundance[sc.recList[q]] *= -1;
	}
	co.sync();
	// ---- coverage filter -d (.cpp:2423-2481) ----
	for (int i = 0; i < T; ++i) {
		if (!(pr.abundance[i] >= 0)) continue;
		int len = len_from_fpkm(prm, pr.abundance[i], pr.fpkm[i]);
		double cov = (pr.abundance[i] * readLen) / len;
		const uint64_t *tb = pv.bits(i);
		int pc = 0;
		for (int w = 0; w < W; ++w) pc += popc64(tb[w]);
		if ((T > 1 || len <= 1000 || pc <= 3) && cov < prm.minDepth) {
			double gmc = sc.gMaxCov[sc.tgid[i] - v.baseGene];
			if (gmc == cov) continue;
			if (gmc <= 5) {
				// is the transcript fully tiled by the constraints compatible with it?
				bool equal = true;
				for (int w = 0; w < W; ++w) {
					uint64_t acc = 0;
					for (int k = co.lane; k < C; k += co.n)
						if (con_in_tx(t, k, tb, pv.first(i), pv.last(i), W)) acc |= t.bits[(long long)k * W + w];
					acc = co.bor(acc);
					if (acc != tb[w]) equal = false;
				}
				co.sync();
				if (equal && co.lane == 0) pr.abundance[i] = -pr.abundance[i];
			} else if (co.lane == 0)
				pr.abundance[i] = -pr.abundance[i];
			co.sync();
		}
	}
	co.sync();
	// ---- too short (.cpp:2486-2496) + chain-support recovery (.cpp:2508-2706) ----
	if (co.lane == 0) {
		for (int i = 0; i < T; ++i) {
			if (pr.abundance[i] <= 0) continue;
			int len = len_from_fpkm(prm, pr.abundance[i], pr.fpkm[i]);
			if (len < 200) pr.abundance[i] = -pr.abundance[i];
		}
		int S = prm.S;
		int nrec = 0;
		int e0 = v.noff[0], E = v.noff[v.n] - e0;
		for (int e = 0; e <= E; ++e) sc.missing[e] = 0;
		for (int i = 0; i < T; ++i) sc.used[i] = 0;
		for (int i = 0; i < v.n && S > 1; ++i) {
			// map order: key {0} (the cs00 quirk, only for i == 0) first, then the edges in next order
			int ne = v.noff[i + 1] - v.noff[i];
			for (int q = (i == 0 ? -1 : 0); q < ne; ++q) {
				int second, b, slot;
				if (q < 0) {
					second = sv.cs00[v.c];
					b = 0;
					slot = E;
					if (second == 0) continue;
				} else {
					int e = v.noff[i] + q;
					second = sv.edge_support[e];
					b = v.nidx[e];
					slot = e - e0;
					if (second == 0) continue;
				}
				if (S >= 0 && (second < 3 || second < (int)(0.5 * S)) && second <= S / 2) continue;
				bool recover = true;
				int j;
				for (j = 0; j < T; ++j) { // Q5: the inner loop reuses j
					if (pr.abundance[j] < 0) continue;
					if (chain_in_tx(pv.bits(j), pv.first(j), pv.last(j), i, b, W)) { recover = false; break; }
					if (recover) {
						for (j = 0; j < T; ++j) {
							if (pr.abundance[j] > 0) continue;
							if (chain_in_tx(pv.bits(j), pv.first(j), pv.last(j), i, b, W)) {
								sc.missing[slot] = 1;
								if (!sc.used[j]) { sc.recList[nrec++] = j; sc.used[j] = 1; }
							}
						}
					}
				}
			}
		}
		for (int i = 0; i < T; ++i) sc.used[i] = 0;
		const int *sup2 = pv.cd->sup2;
		for (int it = 0; it < nrec; ++it) {
			int maxTag = -1;
			for (int j = 0; j < nrec; ++j) {
				int cj = sc.recList[j];
				if (sc.used[cj]) continue;
				if (maxTag == -1) { maxTag = cj; continue; }
				int sj = sup2[pr.id[cj]], sm = sup2[pr.id[maxTag]];
				if (sj > sm) maxTag = cj;
				else if (sj == sm && pr.fpkm[cj] > pr.fpkm[maxTag]) maxTag = cj;
			}
			if (maxTag == -1 || sup2[pr.id[maxTag]] <= 2 || sup2[pr.id[maxTag]] < 0.5 * S) break;
			sc.used[maxTag] = 1;
			int gi = sc.tgid[maxTag] - v.baseGene;
			if (sc.gRecover[gi] >= 1) continue;
			++sc.gRecover[gi];
			int cnt = 0;
			const uint64_t *tb = pv.bits(maxTag);
			int tf = pv.first(maxTag), tl = pv.last(maxTag);
			for (int i = 0; i < v.n; ++i)
				for (int e = v.noff[i]; e < v.noff[i + 1]; ++e)
					if (sc.missing[e - e0] && chain_in_tx(tb, tf, tl, i, v.nidx[e], W)) { sc.missing[e - e0] = 0; ++cnt; }
			if (sc.missing[E] && chain_in_tx(tb, tf, tl, 0, 0, W)) { sc.missing[E] = 0; ++cnt; }
			int len = len_from_fpkm(prm, pr.abundance[maxTag], pr.fpkm[maxTag]);
			double cov = (pr.abundance[maxTag] * readLen) / len;
			if (cnt >= 1 && cov > 1.0) pr.abundance[maxTag] *= -1;
		}
	}
	T = remove_negative(co, pv);
	// ---- transcripts whose support is almost never unique (.cpp:2723-2743) ----
	for (int i = 0; i < T; ++i) {
		const uint64_t *tb = pv.bits(i);
		int tf = pv.first(i), tl = pv.last(i);
		int support = 0, uniq = 0;
		for (int j = co.lane; j < M; j += co.n) {
			if (!con_in_tx(t, t.pi[j], tb, tf, tl, W) || !con_in_tx(t, t.pj[j], tb, tf, tl, W)) continue;
			support += t.psup[j];
			uniq += t.puniq[j];
		}
		support = co.sum(support);
		uniq = co.sum(uniq);
		co.sync();
		if (co.lane == 0 && (double)uniq < 0.03 * support) pr.abundance[i] = -1;
	}
	T = remove_negative(co, pv);
	// ---- shadow two-exon transcripts (.cpp:2749-2834) ----
	int mainStrand;
	{
		int plus = 0, minus = 0;
		for (int i = 0; i < v.n; ++i) {
			if (v.rs[i] == 1) ++plus;
			else if (v.rs[i] == -1) ++minus;
		}
		mainStrand = plus > minus ? 1 : -1;
	}
	for (int i = 0; i < T; ++i) {
		const uint64_t *tb = pv.bits(i);
		int tf = pv.first(i), tl = pv.last(i);
		int intronCnt = 0, anchorA = -1, anchorB = -1, prev = -1;
		int lenA = 0, lenB = 0;
		for (int q = tf; q <= tl; ++q) {
			if (!bt_test(tb, q)) continue;
			if (prev >= 0 && v.en[prev] + 1 < v.st[q]) {
				++intronCnt;
				anchorA = prev;
				anchorB = q;
			}
			prev = q;
		}
		if (intronCnt != 1) continue;
		for (int q = tf; q <= tl; ++q) {
			if (!bt_test(tb, q)) continue;
			if (q <= anchorA) lenA += v.en[q] - v.st[q] + 1;
			else lenB += v.en[q] - v.st[q] + 1;
		}
		int flag = 0;
		if (v.rs[anchorA] == mainStrand) {
			int j = anchorA;
			if (v.en[j] - v.st[j] + 1 <= 20 ||
			    (v.st[j + 1] == v.en[j] + 1 && v.en[j + 1] - v.st[j + 1] + 1 <= 20 && v.rs[j + 1] == mainStrand))
				++flag;
			j = anchorB;
			if (v.en[j] - v.st[j] + 1 <= 20 ||
			    (v.st[j] == v.en[j - 1] + 1 && v.en[j - 1] - v.st[j - 1] + 1 <= 20 && v.ls[j - 1] == mainStrand))
				++flag;
		}
		if (lenA <= 25 || lenB <= 25) flag = 2;
		int support = 0, uniq = 0;
		for (int j = co.lane; j < M; j += co.n) {
			int a = t.pi[j], b = t.pj[j];
			if (!con_in_tx(t, a, tb, tf, tl, W) || !con_in_tx(t, b, tb, tf, tl, W)) continue;
			const uint64_t *ba = t.bits + (long long)a * W, *bb = t.bits + (long long)b * W;
			if ((bt_test(ba, anchorA) && bt_test(ba, anchorB)) || (bt_test(bb, anchorA) && bt_test(bb, anchorB))) {
				support += t.psup[j];
				uniq += t.puniq[j];
			}
		}
		support = co.sum(support);
		uniq = co.sum(uniq);
		if ((double)uniq < 0.3 * support || support < prm.minDepth) flag = 2;
		co.sync();
		if (flag == 2 && co.lane == 0) pr.abundance[i] = -1;
	}
	T = remove_negative(co, pv);
	return T;
}

// =============================================================================================
// phase drivers, one cooperative group per (component, sample) unit
// =============================================================================================
struct UnitCtx {
	CompView v;
	CandView cd;
	TableView t;
	UnitScratch sc;
	UnitPred pr;
	int cap, ul; // ul: unit index inside the chunk
};
PSC_DEV bool unit_setup(const GraphDev &g, const SolveDev &sv, const ChunkDev &ch, int c, int s, UnitCtx &u)
{
	u.v = comp_view(g, c);
	u.cd = cand_view(sv, c, u.v.W);
	u.t = table_view(sv.samples[s], c);
	int A = u.cd.A, M = u.t.M;
	u.cap = A < M + 1 ? A : M + 1;
	if (u.cap < 1) u.cap = 1;
	u.ul = (c - ch.c0) * sv.prm.S + s;
	int E = u.v.noff[u.v.n] - u.v.noff[0];
	unit_scratch_layout(&u.sc, ch.scratch + ch.scratch_off[u.ul], A, M, u.cap, u.v.n, u.t.C, u.v.usedGene - u.v.baseGene, E);
	unit_pred_layout(&u.pr, ch.pred + ch.pred_off[u.ul], u.cap);
	return true;
}

// P3: first round of PickTranscripts; counts how many samples picked each candidate (.cpp:3274-3305)
PSC_DEV void unit_pick_round1(const Coop &co, const GraphDev &g, const SolveDev &sv, const ChunkDev &ch, int c, int s)
{
	UnitCtx u;
	unit_setup(g, sv, ch, c, s, u);
	int T = pick_transcripts(co, u.v, u.cd, u.t, sv.prm, u.sc, u.cap, false, u.pr);
	if (T < 0) { if (co.lane == 0) *sv.error_flag = 12; T = 0; }
	for (int k = co.lane; k < T; k += co.n) atomic_add(&u.cd.sup1[u.pr.id[k]], 1);
}

// P4: second round + filtration (.cpp:3307-3353)
PSC_DEV void unit_pick_round2(const Coop &co, const GraphDev &g, const SolveDev &sv, const ChunkDev &ch, int c, int s)
{
	UnitCtx u;
	unit_setup(g, sv, ch, c, s, u);
	int T = pick_transcripts(co, u.v, u.cd, u.t, sv.prm, u.sc, u.cap, true, u.pr);
	if (T < 0) { if (co.lane == 0) *sv.error_flag = 12; T = 0; }
	co.sync();
	for (int k = co.lane; k < T; k += co.n) {
		u.pr.raw_id[k] = u.pr.id[k];
		u.pr.raw_abundance[k] = u.pr.abundance[k];
		int id = u.pr.id[k];
		int len = tx_len(u.v, u.cd.bits + (long long)id * u.v.W, u.cd.first[id], u.cd.last[id]);
		u.pr.fpkm[k] = to_fpkm(sv.prm, u.pr.abundance[k], len);
	}
	co.sync();
	PredView pv;
	pv.T = T;
	pv.pr = u.pr;
	pv.cd = &u.cd;
	// txptSampleSupport still holds the round-1 counts here (.cpp:3344)
	CandView cd1 = u.cd;
	cd1.sup2 = u.cd.sup1;
	pv.cd = &cd1;
	int size = refine_transcripts(co, u.v, u.t, sv.prm, sv, u.sc, pv, false);
	abundance_em(co, u.v, u.t, sv.prm, u.sc, pv);
	co.sync();
	for (int k = co.lane; k < size; k += co.n) {
		int id = u.pr.id[k];
		int len = tx_len(u.v, u.cd.bits + (long long)id * u.v.W, u.cd.first[id], u.cd.last[id]);
		u.pr.fpkm[k] = to_fpkm(sv.prm, u.pr.abundance[k], len);
	}
	co.sync();
	size = refine_transcripts(co, u.v, u.t, sv.prm, sv, u.sc, pv, true);
	if (co.lane == 0) {
		ch.n_pred[u.ul] = size;
		ch.n_raw[u.ul] = T;
	}
	for (int k = co.lane; k < size; k += co.n) atomic_add(&u.cd.sup2[u.pr.id[k]], 1);
}

// P5: rescue of candidates that survived in many other samples (.cpp:3355-3397) and exon counting
PSC_DEV void unit_rescue(const Coop &co, const GraphDev &g, const SolveDev &sv, const ChunkDev &ch, int c, int s)
{
	UnitCtx u;
	unit_setup(g, sv, ch, c, s, u);
	int psize = ch.n_pred[u.ul], rsize = ch.n_raw[u.ul];
	int S = sv.prm.S;
	int T = psize;
	if (psize != rsize) {
		if (co.lane == 0) {
			for (int j = 0; j < rsize; ++j) {
				int id = u.pr.raw_id[j];
				bool predicted = false;
				for (int k = 0; k < psize; ++k)
					if (u.pr.id[k] == id) { predicted = true; break; }
				int sup = u.cd.sup2[id];
				if (!predicted && (sup >= 3 && sup >= 0.25 * S)) {
					u.pr.id[T] = id;
					u.pr.abundance[T] = u.pr.raw_abundance[j];
					u.pr.fpkm[T] = (double)u.cd.sup1[id]; // nt = alltranscripts[id]: FPKM still holds the round-1 count
					u.pr.corr[T] = -1;
					++T;
				}
			}
		}
		T = co.bcast(T, 0);
		co.sync();
		if (T != psize) {
			PredView pv;
			pv.T = T;
			pv.pr = u.pr;
			pv.cd = &u.cd;
			abundance_em(co, u.v, u.t, sv.prm, u.sc, pv);
		}
	}
	// exon count of every final transcript
	int ne = 0;
	for (int k = co.lane; k < T; k += co.n) {
		int id = u.pr.id[k];
		const uint64_t *tb = u.cd.bits + (long long)id * u.v.W;
		int prev = -1;
		for (int q = u.cd.first[id]; q <= u.cd.last[id]; ++q) {
			if (!bt_test(tb, q)) continue;
			if (prev < 0 || u.v.st[q] != u.v.en[prev] + 1) ++ne;
			prev = q;
		}
	}
	ne = co.sum(ne);
	if (co.lane == 0) {
		ch.n_pred[u.ul] = T;
		ch.n_exon[u.ul] = ne;
	}
}

// output arrays (psc_transcripts layout) on the device
struct OutDev {
	int *comp, *sample, *chr, *gene;
	int8_t *strand;
	long long *exon_off;
	int *exon_start, *exon_end;
	double *fpkm, *cov;
};

// P6: OutputTranscript (TranscriptDecider.cpp:3-111)
PSC_DEV void unit_emit(const Coop &co, const GraphDev &g, const SolveDev &sv, const ChunkDev &ch, int c, int s,
                       const OutDev &o, long long tx0, long long ex0)
{
	UnitCtx u;
	unit_setup(g, sv, ch, c, s, u);
	if (co.lane != 0) return;
	int T = ch.n_pred[u.ul];
	const int *defGene = sv.def_gene + 2 * c;
	long long ex = ex0;
	for (int k = 0; k < T; ++k) {
		int id = u.pr.id[k];
		const uint64_t *tb = u.cd.bits + (long long)id * u.v.W;
		int tf = u.cd.first[id], tl = u.cd.last[id];
		char strand = '.';
		int prev = -1, len = 0;
		long long exStart = ex;
		bool strandSet = false;
		int gid = -1;
		bool gidSet = false;
		for (int q = tf; q <= tl; ++q) {
			if (!bt_test(tb, q)) continue;
			if (!gidSet && u.v.gene[q] != -2) { gid = u.v.gene[q]; gidSet = true; }
			if (prev >= 0 && u.v.en[prev] + 1 < u.v.st[q] && !strandSet) {
				if (u.v.rs[prev] == 1) strand = '+';
				else if (u.v.rs[prev] == -1) strand = '-';
				strandSet = true;
			}
			if (prev >= 0 && u.v.st[q] == u.v.en[prev] + 1)
				o.exon_end[ex - 1] = u.v.en[q] + 1;
			else {
				o.exon_start[ex] = u.v.st[q] + 1;
				o.exon_end[ex] = u.v.en[q] + 1;
				++ex;
			}
			len += u.v.en[q] - u.v.st[q] + 1;
			prev = q;
		}
		if (!gidSet) { // every subexon is shared between genes: fall back on the first intron's strand
			gid = defGene[0];
			int p2 = -1;
			for (int q = tf; q <= tl; ++q) {
				if (!bt_test(tb, q)) continue;
				if (p2 >= 0 && u.v.en[p2] + 1 < u.v.st[q]) { gid = defGene[(u.v.rs[p2] + 1) / 2]; break; }
				p2 = q;
			}
		}
		long long r = tx0 + k;
		o.comp[r] = c;
		o.sample[r] = s;
		o.chr[r] = g.comp_chr[c];
		o.gene[r] = gid;
		o.strand[r] = (int8_t)strand;
		o.exon_off[r] = exStart;
		o.fpkm[r] = u.pr.fpkm[k];
		o.cov[r] = u.pr.abundance[k] * sv.prm.readLen / len;
	}
}

} // namespace pscdev
