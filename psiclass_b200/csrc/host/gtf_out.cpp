#include "gtf_out.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace psc {

static int cmp_tx(const OutTx &a, const OutTx &b)
{ // CompTranscripts (TranscriptDecider.hpp:114-131)
	if (a.gene_id != b.gene_id) return a.gene_id - b.gene_id;
	if (a.exons.size() != b.exons.size()) return (int)a.exons.size() - (int)b.exons.size();
	for (size_t i = 0; i < a.exons.size(); ++i) {
		if (a.exons[i].first != b.exons[i].first) return a.exons[i].first - b.exons[i].first;
		if (a.exons[i].second != b.exons[i].second) return a.exons[i].second - b.exons[i].second;
	}
	return 0;
}

bool write_gtfs(std::vector<OutTx> &q, const std::vector<double> &totalReadCnt, const std::vector<std::string> &chr,
                const std::string &prefix, int argc, char **argv, const std::vector<std::string> &bams, std::string &err,
                std::vector<VoteTx> *forVote)
{
	size_t S = bams.size();
	// ComputeFPKMTPM: per-sample sums run in queue order
	std::vector<double> totalFPK(S, 0.0);
	for (auto &t : q) totalFPK[(size_t)t.sample] += t.fpkm;
	for (auto &t : q) {
		t.tpm = t.fpkm / (totalFPK[(size_t)t.sample] / 1000000.0);
		t.fpkm /= (totalReadCnt[(size_t)t.sample] / 1000000.0);
	}
	std::vector<size_t> order(q.size());
	for (size_t i = 0; i < order.size(); ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
		int c = cmp_tx(q[a], q[b]);
		if (c != 0) return c < 0;
		if (q[a].sample != q[b].sample) return q[a].sample < q[b].sample;
		return a < b;
	});
	std::vector<FILE *> fps(S, nullptr);
	for (size_t s = 0; s < S; ++s) {
		std::string path = prefix.empty() ? "sample_" + std::to_string(s) + ".gtf" : prefix + "_sample_" + std::to_string(s) + ".gtf";
		fps[s] = fopen(path.c_str(), "w");
		if (!fps[s]) {
			err = "cannot write " + path;
			return false;
		}
		fprintf(fps[s], "#PsiCLASS_v1.0.3\n#");
		for (int a = 0; a < argc; ++a) fprintf(fps[s], "%s%s", argv[a], a + 1 < argc ? " " : "\n");
		const char *base = strrchr(bams[s].c_str(), '/');
		fprintf(fps[s], "#%s\n", base ? base + 1 : bams[s].c_str());
	}
	std::vector<std::vector<VoteTx>> perSample(forVote ? S : 0);
	int gid = -1, tid = 0;
	for (size_t i = 0; i < order.size();) {
		size_t j = i + 1;
		while (j < order.size() && cmp_tx(q[order[i]], q[order[j]]) == 0) ++j;
		if (q[order[i]].gene_id != gid) {
			gid = q[order[i]].gene_id;
			tid = 0;
		} else
			++tid;
		for (size_t l = i; l < j; ++l) {
			const OutTx &t = q[order[l]];
			const char *c = chr[(size_t)t.chr].c_str();
			FILE *fp = fps[(size_t)t.sample];
			fprintf(fp, "%s\tPsiCLASS\ttranscript\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\";\n",
			        c, t.exons[0].first, t.exons.back().second, t.strand, c, t.gene_id, c, t.gene_id, tid, t.fpkm, t.tpm, t.cov);
			for (size_t e = 0; e < t.exons.size(); ++e)
				fprintf(fp, "%s\tPsiCLASS\texon\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; exon_number \"%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\";\n",
				        c, t.exons[e].first, t.exons[e].second, t.strand, c, t.gene_id, c, t.gene_id, tid, (int)e + 1, t.fpkm, t.tpm, t.cov);
			if (forVote) {
				VoteTx v;
				v.sample = t.sample;
				v.gene_id = t.gene_id;
				v.transcript_id = tid;
				v.chrom = c;
				v.strand = t.strand;
				v.exons = t.exons;
				v.fpkm = gtf_rounded(t.fpkm);
				v.tpm = gtf_rounded(t.tpm);
				v.cov = gtf_rounded(t.cov);
				perSample[(size_t)t.sample].push_back(std::move(v));
			}
		}
		i = j;
	}
	for (FILE *fp : fps) fclose(fp);
	if (forVote)
		for (auto &ps : perSample)
			for (auto &v : ps) forVote->push_back(std::move(v));
	return true;
}

} // namespace psc
