// psc_oracle -- command-line front end of the CPU restatement (TEST INFRASTRUCTURE).
// Same argv as the reference's `classes` (classes.cpp:12-41); writes <prefix>_sample_<i>.gtf.
// With --dump it also prints the tables ref_dump prints, in the same format, so the two can be
// diffed byte for byte.  File parsing (BAM, subexon file) is shared with the product's host
// code and is pinned separately against ref_dump's L/C/e/T lines.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../psiclass_b200/csrc/host/bam_io.h"
#include "../psiclass_b200/csrc/host/subexon_graph.h"
#include "psc_oracle.h"

using namespace orc;

static void graph_component(const psc::Graph &g, int c, Component &out)
{
	out = Component();
	out.chr = g.comp_chr[c];
	out.baseGeneId = g.comp_base_gene[c];
	out.usedGeneId = g.comp_used_gene[c];
	int b = g.comp_off[c], e = g.comp_off[c + 1];
	out.se.resize((size_t)(e - b));
	for (int i = b; i < e; ++i) {
		Subexon &s = out.se[(size_t)(i - b)];
		s.start = g.se_start[i];
		s.end = g.se_end[i];
		s.lt = g.se_left_type[i];
		s.rt = g.se_right_type[i];
		s.ls = g.se_left_strand[i];
		s.rs = g.se_right_strand[i];
		s.geneId = g.se_gene_id[i];
		s.lcls = g.se_left_cls[i];
		s.rcls = g.se_right_cls[i];
		s.next.assign(g.next_idx.begin() + g.next_off[i], g.next_idx.begin() + g.next_off[i + 1]);
		s.prev.assign(g.prev_idx.begin() + g.prev_off[i], g.prev_idx.begin() + g.prev_off[i + 1]);
	}
}

static void reads_view(const psc::SampleReads &s, Reads &r)
{
	r = Reads();
	r.chr_off = s.chr_off;
	r.pos = s.pos.data();
	r.mate_pos = s.mate_pos.data();
	r.flags = s.flags.data();
	r.nh = s.nh.data();
	r.name_key = s.name_key.data();
	r.seg_off = s.seg_off.data();
	r.seg_a = s.seg_a.data();
	r.seg_b = s.seg_b.data();
	r.n = s.n_reads();
	r.read_len = s.read_len;
	r.frag_len = s.frag_len;
	r.frag_stdev = s.frag_stdev;
	r.mate_paired = s.mate_paired;
}

static int cmp_tx(const OutTranscript &a, const OutTranscript &b)
{ // MultiThreadOutputTranscript::CompTranscripts (TranscriptDecider.hpp:114-131)
	if (a.geneId != b.geneId) return a.geneId - b.geneId;
	if (a.exons.size() != b.exons.size()) return (int)a.exons.size() - (int)b.exons.size();
	for (size_t i = 0; i < a.exons.size(); ++i) {
		if (a.exons[i].first != b.exons[i].first) return a.exons[i].first - b.exons[i].first;
		if (a.exons[i].second != b.exons[i].second) return a.exons[i].second - b.exons[i].second;
	}
	return 0;
}

int main(int argc, char **argv)
{
	std::string subexonPath, prefix;
	std::vector<std::string> bamPaths;
	Params prm;
	bool dump = false, tables = true;
	for (int a = 1; a < argc; ++a) {
		std::string o = argv[a];
		if (o == "-s") subexonPath = argv[++a];
		else if (o == "-b") bamPaths.push_back(argv[++a]);
		else if (o == "--lb") {
			FILE *fp = fopen(argv[++a], "r");
			if (!fp) { printf("Cannot open %s\n", argv[a]); return 1; }
			char line[1024];
			while (fgets(line, sizeof(line), fp)) {
				size_t n = strlen(line);
				while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
				if (n) bamPaths.push_back(line);
			}
			fclose(fp);
		} else if (o == "-o") prefix = argv[++a];
		else if (o == "-p") ++a;
		else if (o == "-c") prm.classifierThreshold = atof(argv[++a]);
		else if (o == "-f") prm.fpkmFraction = atof(argv[++a]);
		else if (o == "-d") prm.minDepth = atof(argv[++a]);
		else if (o == "--maxDpConstraintSize") prm.maxDpConstraintSize = atoi(argv[++a]);
		else if (o == "--primaryParalog") prm.primaryAsUnique = true;
		else if (o == "--hasMateIdSuffix") ;
		else if (o == "--stranded") ++a;
		else if (o == "--dump") dump = true;
		else if (o == "--no-tables") tables = false;
		else { printf("unknown option %s\n", argv[a]); return 1; }
	}
	if (subexonPath.empty() || bamPaths.empty()) { printf("need -s and -b/--lb\n"); return 1; }
	int S = (int)bamPaths.size();
	std::vector<psc::SampleReads> samples((size_t)S);
	std::string err;
	for (int s = 0; s < S; ++s)
		if (!psc::decode_bam(bamPaths[(size_t)s], 4, samples[(size_t)s], err)) {
			fprintf(stderr, "%s\n", err.c_str());
			return 1;
		}
	if (dump)
		for (int s = 0; s < S; ++s)
			printf("L %d %d %d %d %d\n", s, samples[s].read_len, samples[s].frag_len, samples[s].frag_stdev,
			       samples[s].mate_paired);
	psc::Graph g;
	if (!psc::load_subexon_graph(subexonPath, samples[0].chr_names, prm.classifierThreshold, g, err)) {
		printf("%s\n", err.c_str());
		return 1;
	}
	prm.sampleCnt = S;
	prm.readLen = samples[0].read_len;
	prm.fragLen = samples[0].frag_len;
	prm.fragStdev = samples[0].frag_stdev;
	prm.matePaired = samples[0].mate_paired != 0;
	std::vector<Reads> reads((size_t)S);
	for (int s = 0; s < S; ++s) reads_view(samples[s], reads[s]);
	if (dump) printf("G %d %d\n", g.n_comp, g.n_se());

	std::vector<OutTranscript> out;
	std::vector<ConstraintSet> cs((size_t)S);
	for (int c = 0; c < g.n_comp; ++c) {
		Component comp;
		graph_component(g, c, comp);
		int n = (int)comp.se.size();
		if (dump) {
			printf("C %d %d %d %d %d %d %d\n", c, n, comp.chr, g.comp_start[c], g.comp_end[c], comp.baseGeneId,
			       comp.usedGeneId);
			if (tables)
				for (int k = 0; k < n; ++k) {
					const Subexon &s = comp.se[k];
					printf("e %d %d %d %d %d %d %d %d %.17g %.17g %d", k, s.start, s.end, s.lt, s.rt, s.ls, s.rs,
					       s.geneId, s.lcls, s.rcls, (int)s.prev.size());
					for (int t : s.prev) printf(" %d", t);
					printf(" %d", (int)s.next.size());
					for (int t : s.next) printf(" %d", t);
					printf("\n");
				}
		} else
			printf("%d: %d %s %d %d\n", c, n, g.chr_names[comp.chr].c_str(), g.comp_start[c] + 1, g.comp_end[c] + 1);
		for (int s = 0; s < S; ++s) {
			build_constraints(comp, reads[s], prm.primaryAsUnique, cs[s]);
			if (dump && tables) {
				printf("S %d %d %d %d\n", c, s, (int)cs[s].c.size(), (int)cs[s].p.size());
				std::vector<int> ones;
				for (size_t k = 0; k < cs[s].c.size(); ++k) {
					const Constraint &q = cs[s].c[k];
					printf("c %d %d %d %d %d %d %.17g %.17g", (int)k, q.first, q.last, q.support, q.uniq,
					       q.maxReadLen, q.weight, q.normAbund);
					ones.clear();
					for (int b = 0; b < n; ++b)
						if (bit_test(q.v, b)) ones.push_back(b);
					printf(" %d", (int)ones.size());
					for (int b : ones) printf(" %d", b);
					printf("\n");
				}
				for (size_t k = 0; k < cs[s].p.size(); ++k) {
					const MatePair &m = cs[s].p[k];
					printf("m %d %d %d %d %d %d %.17g\n", (int)k, m.i, m.j, m.support, m.uniq, m.effectiveCount,
					       m.normAbund);
				}
			}
		}
		SolveStats st;
		int c0 = (int)cs[0].c.size(), m0 = (int)cs[0].p.size();
		solve(c, comp, cs, prm, out, &st);
		printf("%d: atCnt=%d seCnt=%d %d %d %d\n", comp.se[0].start + 1, st.atCnt, n, st.useDP, c0, m0);
	}
	// Q8: totalReadCnt = #primary records read when the stream stopped
	std::vector<double> totalReadCnt((size_t)S);
	for (int s = 0; s < S; ++s) {
		int64_t lr = reads[s].last_read;
		uint32_t t = (lr < 0) ? 0 : (lr >= reads[s].n ? samples[s].total_primary : samples[s].prim_cum[(size_t)lr]);
		totalReadCnt[s] = t;
		if (dump) printf("T %d %u\n", s, t);
	}
	// ComputeFPKMTPM + Flush (TranscriptDecider.hpp:190-283)
	std::vector<double> totalFPK((size_t)S, 0.0);
	for (auto &t : out) totalFPK[t.sample] += t.fpkm;
	std::vector<double> tpm(out.size());
	for (size_t i = 0; i < out.size(); ++i) {
		tpm[i] = out[i].fpkm / (totalFPK[out[i].sample] / 1000000.0);
		out[i].fpkm /= (totalReadCnt[out[i].sample] / 1000000.0);
	}
	std::vector<size_t> order(out.size());
	for (size_t i = 0; i < order.size(); ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
		int c = cmp_tx(out[a], out[b]);
		if (c != 0) return c < 0;
		if (out[a].sample != out[b].sample) return out[a].sample < out[b].sample;
		return a < b;
	});
	std::vector<FILE *> fps((size_t)S);
	for (int s = 0; s < S; ++s) {
		std::string path = prefix.empty() ? "sample_" + std::to_string(s) + ".gtf"
		                                  : prefix + "_sample_" + std::to_string(s) + ".gtf";
		fps[s] = fopen(path.c_str(), "w");
		fprintf(fps[s], "#PsiCLASS_v1.0.3\n#");
		for (int a = 0; a < argc; ++a) fprintf(fps[s], "%s%s", argv[a], a + 1 < argc ? " " : "\n");
		const char *base = strrchr(bamPaths[s].c_str(), '/');
		fprintf(fps[s], "#%s\n", base ? base + 1 : bamPaths[s].c_str());
	}
	int gid = -1, tid = 0;
	for (size_t i = 0; i < order.size();) {
		size_t j = i + 1;
		while (j < order.size() && cmp_tx(out[order[i]], out[order[j]]) == 0) ++j;
		if (out[order[i]].geneId != gid) {
			gid = out[order[i]].geneId;
			tid = 0;
		} else
			++tid;
		for (size_t l = i; l < j; ++l) {
			const OutTranscript &t = out[order[l]];
			const char *chrom = g.chr_names[t.chr].c_str();
			FILE *fp = fps[t.sample];
			fprintf(fp, "%s\tPsiCLASS\ttranscript\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\";\n",
			        chrom, t.exons[0].first, t.exons.back().second, t.strand, chrom, t.geneId, chrom, t.geneId, tid,
			        t.fpkm, tpm[order[l]], t.cov);
			for (size_t e = 0; e < t.exons.size(); ++e)
				fprintf(fp, "%s\tPsiCLASS\texon\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; exon_number \"%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\";\n",
				        chrom, t.exons[e].first, t.exons[e].second, t.strand, chrom, t.geneId, chrom, t.geneId, tid,
				        (int)e + 1, t.fpkm, tpm[order[l]], t.cov);
		}
		i = j;
	}
	for (FILE *fp : fps) fclose(fp);
	return 0;
}
