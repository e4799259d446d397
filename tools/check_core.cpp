// check_core -- TEST TOOL (shared by psc_check and the pytest testkit): runs the engine behind the C ABI (libpsc.so on a GPU box, or the CPU
// shim build under tests/shim/_build when debugging without a GPU) and the CPU oracle
// (oracle/psc_oracle.cpp) on the same dataset and compares every table:
//   constraint tables (bit-exact), atCnt/useDP, and the final transcripts (structures exact,
//   fpkm/cov within 1e-9 relative; the number that are bitwise equal is reported too).
// usage: psc_check <dir with subexon_combined.out + bamlist> [--maxDp n] [--coop 1|32] [--primaryParalog]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/psc.h"
#include "../oracle/psc_oracle.h"
#include "../psiclass_b200/csrc/host/bam_io.h"
#include "../psiclass_b200/csrc/host/subexon_graph.h"

static void graph_component(const psc::Graph &g, int c, orc::Component &out)
{
	out = orc::Component();
	out.chr = g.comp_chr[c];
	out.baseGeneId = g.comp_base_gene[c];
	out.usedGeneId = g.comp_used_gene[c];
	int b = g.comp_off[c], e = g.comp_off[c + 1];
	out.se.resize((size_t)(e - b));
	for (int i = b; i < e; ++i) {
		orc::Subexon &s = out.se[(size_t)(i - b)];
		s.start = g.se_start[i]; s.end = g.se_end[i];
		s.lt = g.se_left_type[i]; s.rt = g.se_right_type[i];
		s.ls = g.se_left_strand[i]; s.rs = g.se_right_strand[i];
		s.geneId = g.se_gene_id[i];
		s.lcls = g.se_left_cls[i]; s.rcls = g.se_right_cls[i];
		s.next.assign(g.next_idx.begin() + g.next_off[i], g.next_idx.begin() + g.next_off[i + 1]);
		s.prev.assign(g.prev_idx.begin() + g.prev_off[i], g.prev_idx.begin() + g.prev_off[i + 1]);
	}
}


#include "check_core.h"

int run_check(const psc::Graph &g, std::vector<psc::SampleReads> &samples, const EngineApi &api, const CheckOpts &opt, CheckReport &rep)
{
	int S = (int)samples.size();
	int maxDp = opt.maxDp, coop = opt.coop, verbose = opt.verbose;
	bool pau = opt.primaryParalog != 0, skipSolve = opt.constraintsOnly != 0;
	memset(&rep, 0, sizeof(rep));
#define psc_create api.create
#define psc_destroy api.destroy
#define psc_last_error api.last_error
#define psc_set_params api.set_params
#define psc_set_graph api.set_graph
#define psc_set_reads api.set_reads
#define psc_build_constraints api.build_constraints
#define psc_constraint_counts api.constraint_counts
#define psc_get_constraint_table api.get_constraint_table
#define psc_solve api.solve
#define psc_transcript_counts api.transcript_counts
#define psc_get_transcripts api.get_transcripts
#define psc_get_comp_stats api.get_comp_stats
#define psc_get_timing api.get_timing
#define psc_set_coop_width api.set_coop_width
	// ---------------- engine ----------------
	psc_ctx *ctx = nullptr;
	int rc = psc_create(&ctx, opt.device);
	if (rc) { fprintf(stderr, "psc_create failed: %d\n", rc); rep.pass = 0; return 3; }
	psc_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.fpkm_fraction = 0.05; prm.classifier_threshold = 0.05; prm.min_read_depth = 2.5;
	prm.max_dp_constraint_size = maxDp; prm.primary_as_unique = pau; prm.n_samples = S;
	prm.read_len = samples[0].read_len; prm.frag_len = samples[0].frag_len; prm.frag_stdev = samples[0].frag_stdev;
	prm.mate_paired = samples[0].mate_paired;
#define E(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, psc_last_error(ctx)); return 3; } } while (0)
	E(psc_set_params(ctx, &prm));
	E(psc_set_coop_width(ctx, coop));
	psc_graph pg;
	pg.n_comp = g.n_comp; pg.n_se = g.n_se();
	pg.comp_off = g.comp_off.data(); pg.comp_chr = g.comp_chr.data();
	pg.comp_base_gene = g.comp_base_gene.data(); pg.comp_used_gene = g.comp_used_gene.data();
	pg.se_start = g.se_start.data(); pg.se_end = g.se_end.data();
	pg.se_left_type = g.se_left_type.data(); pg.se_right_type = g.se_right_type.data();
	pg.se_left_strand = g.se_left_strand.data(); pg.se_right_strand = g.se_right_strand.data();
	pg.se_left_cls = g.se_left_cls.data(); pg.se_right_cls = g.se_right_cls.data();
	pg.se_gene_id = g.se_gene_id.data();
	pg.next_off = g.next_off.data(); pg.next_idx = g.next_idx.data();
	pg.prev_off = g.prev_off.data(); pg.prev_idx = g.prev_idx.data();
	E(psc_set_graph(ctx, &pg));
	for (int s = 0; s < S; ++s) {
		psc_reads r;
		const psc::SampleReads &sr = samples[s];
		r.n_reads = sr.n_reads(); r.n_segs = (int64_t)sr.seg_a.size(); r.n_chr = (int)sr.chr_names.size();
		r.chr_off = (const int64_t *)sr.chr_off.data(); r.pos = sr.pos.data(); r.mate_pos = sr.mate_pos.data();
		r.flags = sr.flags.data(); r.nh = sr.nh.data(); r.name_key = sr.name_key.data();
		r.seg_off = sr.seg_off.data(); r.seg_a = sr.seg_a.data(); r.seg_b = sr.seg_b.data();
		r.read_len = sr.read_len; r.frag_stdev = sr.frag_stdev;
		E(psc_set_reads(ctx, s, &r));
	}
	E(psc_build_constraints(ctx));
	std::vector<int32_t> nC((size_t)S * g.n_comp), nM((size_t)S * g.n_comp);
	E(psc_constraint_counts(ctx, nC.data(), nM.data()));

	// ---------------- oracle + comparison ----------------
	orc::Params op;
	op.maxDpConstraintSize = maxDp; op.primaryAsUnique = pau; op.sampleCnt = S;
	op.readLen = prm.read_len; op.fragLen = prm.frag_len; op.fragStdev = prm.frag_stdev; op.matePaired = prm.mate_paired != 0;
	std::vector<orc::Reads> reads((size_t)S);
	for (int s = 0; s < S; ++s) {
		orc::Reads &r = reads[s];
		const psc::SampleReads &sr = samples[s];
		r.chr_off = sr.chr_off; r.pos = sr.pos.data(); r.mate_pos = sr.mate_pos.data(); r.flags = sr.flags.data();
		r.nh = sr.nh.data(); r.name_key = sr.name_key.data(); r.seg_off = sr.seg_off.data(); r.seg_a = sr.seg_a.data();
		r.seg_b = sr.seg_b.data(); r.n = sr.n_reads(); r.read_len = sr.read_len; r.frag_len = sr.frag_len;
		r.frag_stdev = sr.frag_stdev; r.mate_paired = sr.mate_paired;
	}
	std::vector<orc::OutTranscript> otx;
	std::vector<orc::SolveStats> ostats((size_t)g.n_comp);
	long long tblBad = 0, tblChecked = 0;
	std::vector<orc::ConstraintSet> cs((size_t)S);
	for (int c = 0; c < g.n_comp; ++c) {
		orc::Component comp;
		graph_component(g, c, comp);
		int n = (int)comp.se.size(), W = (n + 63) / 64;
		for (int s = 0; s < S; ++s) {
			orc::build_constraints(comp, reads[s], pau, cs[s]);
			int C = nC[(size_t)s * g.n_comp + c], M = nM[(size_t)s * g.n_comp + c];
			++tblChecked;
			bool bad = C != (int)cs[s].c.size() || M != (int)cs[s].p.size();
			if (!bad) {
				std::vector<uint64_t> bits((size_t)C * W + 1);
				std::vector<int> first(C + 1), last(C + 1), sup(C + 1), uq(C + 1), ml(C + 1), pi(M + 1), pj(M + 1), ps(M + 1), pu(M + 1), pe(M + 1);
				std::vector<double> wt(C + 1), na(C + 1), pn(M + 1);
				psc_constraint_table t;
				t.bits = bits.data(); t.first = first.data(); t.last = last.data(); t.support = sup.data();
				t.uniq_support = uq.data(); t.max_read_len = ml.data(); t.weight = wt.data(); t.norm_abund = na.data();
				t.pair_i = pi.data(); t.pair_j = pj.data(); t.pair_support = ps.data(); t.pair_uniq = pu.data();
				t.pair_effective_count = pe.data(); t.pair_norm_abund = pn.data();
				E(psc_get_constraint_table(ctx, s, c, &t));
				for (int k = 0; k < C && !bad; ++k) {
					const orc::Constraint &q = cs[s].c[k];
					if (first[k] != q.first || last[k] != q.last || sup[k] != q.support || uq[k] != q.uniq || ml[k] != q.maxReadLen ||
					    wt[k] != q.weight || na[k] != q.normAbund || memcmp(&bits[(size_t)k * W], q.v.data(), 8 * (size_t)W) != 0) {
						bad = true;
						if (verbose) fprintf(stderr, "  constraint %d differs: first %d/%d last %d/%d sup %d/%d uq %d/%d ml %d/%d w %.17g/%.17g na %.17g/%.17g\n", k,
						        first[k], q.first, last[k], q.last, sup[k], q.support, uq[k], q.uniq, ml[k], q.maxReadLen, wt[k], q.weight, na[k], q.normAbund);
					}
				}
				for (int k = 0; k < M && !bad; ++k) {
					const orc::MatePair &m = cs[s].p[k];
					if (pi[k] != m.i || pj[k] != m.j || ps[k] != m.support || pu[k] != m.uniq || pe[k] != m.effectiveCount || pn[k] != m.normAbund) {
						bad = true;
						if (verbose) fprintf(stderr, "  pair %d differs: (%d,%d)/(%d,%d) sup %d/%d uq %d/%d\n", k, pi[k], pj[k], m.i, m.j, ps[k], m.support, pu[k], m.uniq);
					}
				}
			}
			if (bad) {
				++tblBad;
				if (tblBad <= 10) fprintf(stderr, "TABLE MISMATCH comp %d sample %d: engine C=%d M=%d oracle C=%d M=%d\n", c, s, C, M, (int)cs[s].c.size(), (int)cs[s].p.size());
			}
		}
		if (!skipSolve) orc::solve(c, comp, cs, op, otx, &ostats[c]);
	}
	if (verbose >= 0) printf("constraint tables: %lld checked, %lld mismatching\n", tblChecked, tblBad);
	if (skipSolve) { psc_destroy(ctx); rep.n_comp = g.n_comp; rep.tables_checked = tblChecked; rep.tables_bad = tblBad; rep.pass = tblBad ? 0 : 1; return tblBad ? 1 : 0; }

	E(psc_solve(ctx));
	std::vector<int32_t> atc(g.n_comp + 1), udp(g.n_comp + 1), ncd(g.n_comp + 1);
	psc_comp_stats stt; stt.at_cnt = atc.data(); stt.use_dp = udp.data(); stt.n_candidates = ncd.data();
	E(psc_get_comp_stats(ctx, &stt));
	int statBad = 0;
	for (int c = 0; c < g.n_comp; ++c)
		if (atc[c] != ostats[c].atCnt || udp[c] != ostats[c].useDP) {
			if (++statBad <= 10) fprintf(stderr, "STATS MISMATCH comp %d: atCnt %d/%d useDP %d/%d\n", c, atc[c], ostats[c].atCnt, udp[c], ostats[c].useDP);
		}
	int64_t nt = 0, ne = 0;
	E(psc_transcript_counts(ctx, &nt, &ne));
	std::vector<int32_t> tc(nt + 1), ts(nt + 1), tchr(nt + 1), tg(nt + 1), es(ne + 1), ee(ne + 1);
	std::vector<int8_t> tstr(nt + 1);
	std::vector<int64_t> eo(nt + 2);
	std::vector<double> tf(nt + 1), tcov(nt + 1);
	psc_transcripts pt;
	pt.comp = tc.data(); pt.sample = ts.data(); pt.chr = tchr.data(); pt.gene_id = tg.data(); pt.strand = tstr.data();
	pt.exon_off = eo.data(); pt.exon_start = es.data(); pt.exon_end = ee.data(); pt.fpkm = tf.data(); pt.cov = tcov.data();
	E(psc_get_transcripts(ctx, &pt));
	long long txBad = 0, valExact = 0, valClose = 0, valBad = 0;
	if ((size_t)nt != otx.size()) fprintf(stderr, "TRANSCRIPT COUNT MISMATCH: engine %lld oracle %zu\n", (long long)nt, otx.size());
	size_t cmpN = std::min((size_t)nt, otx.size());
	int firstBadComp = -1;
	for (size_t k = 0; k < cmpN; ++k) {
		const orc::OutTranscript &o = otx[k];
		bool same = tc[k] == o.comp && ts[k] == o.sample && tchr[k] == o.chr && tg[k] == o.geneId && (char)tstr[k] == o.strand &&
		            (size_t)(eo[k + 1] - eo[k]) == o.exons.size();
		for (size_t e = 0; same && e < o.exons.size(); ++e)
			same = es[eo[k] + e] == o.exons[e].first && ee[eo[k] + e] == o.exons[e].second;
		if (!same) {
			++txBad;
			if (firstBadComp < 0) firstBadComp = std::min(tc[k], o.comp);
			continue;
		}
		auto rel = [](double a, double b) { return a == b ? 0.0 : fabs(a - b) / std::max(fabs(a), fabs(b)); };
		double r = std::max(rel(tf[k], o.fpkm), rel(tcov[k], o.cov));
		if (r == 0) ++valExact;
		else if (r <= 1e-9) ++valClose;
		else {
			++valBad;
			if (valBad <= 5) fprintf(stderr, "VALUE MISMATCH tx %zu comp %d sample %d: fpkm %.17g/%.17g cov %.17g/%.17g\n", k, tc[k], ts[k], tf[k], o.fpkm, tcov[k], o.cov);
		}
	}
	if (firstBadComp >= 0) fprintf(stderr, "first structural mismatch in component %d (useDP=%d)\n", firstBadComp, udp[firstBadComp]);
	psc_timing tm;
	psc_get_timing(ctx, &tm);
	if (verbose >= 0) printf("components %d (DP %d), stats mismatching %d\n", g.n_comp, (int)std::count(udp.begin(), udp.begin() + g.n_comp, 1), statBad);
	if (verbose >= 0) printf("transcripts: engine %lld oracle %zu, structure mismatches %lld, values exact %lld, within 1e-9 %lld, off %lld\n",
	       (long long)nt, otx.size(), txBad, valExact, valClose, valBad);
	if (verbose >= 0) printf("device ms: constraints %.3f solve %.3f, launches %lld\n", tm.constraints_ms, tm.solve_ms, (long long)tm.kernel_launches);

	psc_destroy(ctx);
	bool fail = tblBad || statBad || txBad || valBad || (size_t)nt != otx.size();
	rep.n_comp = g.n_comp; rep.n_dp = (int)std::count(udp.begin(), udp.begin() + g.n_comp, 1);
	rep.tables_checked = tblChecked; rep.tables_bad = tblBad; rep.stats_bad = statBad;
	rep.tx_engine = nt; rep.tx_oracle = (long long)otx.size(); rep.tx_bad = txBad;
	rep.val_exact = valExact; rep.val_close = valClose; rep.val_bad = valBad;
	rep.constraints_ms = tm.constraints_ms; rep.solve_ms = tm.solve_ms; rep.launches = tm.kernel_launches;
	rep.n_downsampled = rep.n_iter10 = rep.n_skipped = 0;
	for (auto &o : ostats) { rep.n_downsampled += o.nDownsampled; rep.n_iter10 += o.nIter10; rep.n_skipped += o.nSkipped; }
	rep.pass = fail ? 0 : 1;
	return fail ? 1 : 0;
}
