#include "gtf_out.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

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
	// transcript ids (Flush, TranscriptDecider.hpp:236-256) and the per-sample order of the records
	std::vector<int> tidOf(q.size(), 0);
	std::vector<std::vector<size_t>> bySample(S);
	{
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
				tidOf[order[l]] = tid;
				bySample[(size_t)q[order[l]].sample].push_back(order[l]);
			}
			i = j;
		}
	}
	// one writer thread per sample file (the text is ~140 bytes per exon line: formatting dominates)
	std::vector<std::vector<VoteTx>> perSample(forVote ? S : 0);
	std::vector<std::string> errs(S);
	auto writeSample = [&](size_t sIdx) {
		std::string path = prefix.empty() ? "sample_" + std::to_string(sIdx) + ".gtf" : prefix + "_sample_" + std::to_string(sIdx) + ".gtf";
		FILE *fp = fopen(path.c_str(), "w");
		if (!fp) {
			errs[sIdx] = "cannot write " + path;
			return;
		}
		std::vector<char> iobuf(1 << 20);
		setvbuf(fp, iobuf.data(), _IOFBF, iobuf.size());
		fprintf(fp, "#PsiCLASS_v1.0.3\n#");
		for (int a = 0; a < argc; ++a) fprintf(fp, "%s%s", argv[a], a + 1 < argc ? " " : "\n");
		const char *baseName = strrchr(bams[sIdx].c_str(), '/');
		fprintf(fp, "#%s\n", baseName ? baseName + 1 : bams[sIdx].c_str());
		char tail[160];
		for (size_t idx : bySample[sIdx]) {
			const OutTx &t = q[idx];
			const int tid = tidOf[idx];
			const char *c = chr[(size_t)t.chr].c_str();
			// the part every line of the transcript shares
			snprintf(tail, sizeof(tail), "FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\";\n", t.fpkm, t.tpm, t.cov);
			fprintf(fp, "%s\tPsiCLASS\ttranscript\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; %s", c, t.exons[0].first,
			        t.exons.back().second, t.strand, c, t.gene_id, c, t.gene_id, tid, tail);
			for (size_t e = 0; e < t.exons.size(); ++e)
				fprintf(fp, "%s\tPsiCLASS\texon\t%d\t%d\t1000\t%c\t.\tgene_id \"%s.%d\"; transcript_id \"%s.%d.%d\"; exon_number \"%d\"; %s", c,
				        t.exons[e].first, t.exons[e].second, t.strand, c, t.gene_id, c, t.gene_id, tid, (int)e + 1, tail);
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
				perSample[sIdx].push_back(std::move(v));
			}
		}
		fclose(fp);
	};
	{
		size_t nThreads = std::min<size_t>(S, std::max(1u, std::thread::hardware_concurrency()));
		std::vector<std::thread> pool;
		std::atomic<size_t> next{0};
		for (size_t t = 0; t < nThreads; ++t)
			pool.emplace_back([&]() {
				for (;;) {
					size_t sIdx = next.fetch_add(1);
					if (sIdx >= S) break;
					writeSample(sIdx);
				}
			});
		for (auto &th : pool) th.join();
	}
	for (auto &e : errs)
		if (!e.empty()) {
			err = e;
			return false;
		}
	if (forVote)
		for (auto &ps : perSample)
			for (auto &v : ps) forVote->push_back(std::move(v));
	return true;
}

} // namespace psc
