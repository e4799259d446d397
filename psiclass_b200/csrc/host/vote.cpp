#include "vote.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>

namespace psc {

double gtf_rounded(double x)
{
	char buf[400];
	snprintf(buf, sizeof(buf), "%.6lf", x);
	return atof(buf);
}

static int cmp_tx(const VoteTx &a, const VoteTx &b)
{ // CompTranscripts (TranscriptDecider.hpp:114-131): the chromosome is not part of the key (gene ids are global)
	if (a.gene_id != b.gene_id) return a.gene_id - b.gene_id;
	if (a.exons.size() != b.exons.size()) return (int)a.exons.size() - (int)b.exons.size();
	for (size_t i = 0; i < a.exons.size(); ++i) {
		if (a.exons[i].first != b.exons[i].first) return a.exons[i].first - b.exons[i].first;
		if (a.exons[i].second != b.exons[i].second) return a.exons[i].second - b.exons[i].second;
	}
	return 0;
}

void vote_core(std::vector<VoteTx> &txs, int sampleCnt, double minAvgDepth, FILE *out)
{
	const double fraction = 1.0;
	int minSampleCnt = 3;
	// std::sort on the sequence the reference builds, with its comparator: equal elements land where they land there
	std::vector<int> order(txs.size());
	for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
	std::sort(order.begin(), order.end(), [&](int x, int y) {
		int c = cmp_tx(txs[(size_t)x], txs[(size_t)y]);
		if (c != 0) return c < 0;
		return txs[(size_t)x].sample < txs[(size_t)y].sample;
	});
	if (minSampleCnt > sampleCnt) minSampleCnt = sampleCnt;
	struct Kept { int rep; double fpkm, tpm, cov; int support; };
	std::vector<Kept> kept; // Vote.cpp:263-312 (its index k starts at 0 in the shipped binary) and :318-405
	const int size = (int)order.size();
	for (int i = 0; i < size;) {
		int j;
		for (j = i + 1; j < size; ++j)
			if (cmp_tx(txs[(size_t)order[(size_t)i]], txs[(size_t)order[(size_t)j]])) break;
		double sumFPKM = 0, sumTPM = 0, sumCov = 0;
		for (int l = i; l < j; ++l) {
			const VoteTx &t = txs[(size_t)order[(size_t)l]];
			sumFPKM += t.fpkm;
			sumTPM += t.tpm;
			sumCov += t.cov;
		}
		const bool drop = (j - i < (int)(fraction * sampleCnt) || j - i < minSampleCnt) && sumCov < minAvgDepth * sampleCnt;
		if (!drop) kept.push_back(Kept{order[(size_t)i], sumFPKM / (j - i), sumTPM / (j - i), sumCov / (j - i), j - i});
		i = j;
	}
	const char *prefix = ""; // Vote.cpp:199-224 derives it from the cov field, which never has two dots
	for (const Kept &k : kept) {
		const VoteTx &t = txs[(size_t)k.rep];
		const char *chrom = t.chrom.c_str();
		fprintf(out, "%s\tPsiCLASS\ttranscript\t%d\t%d\t1000\t%c\t.\tgene_id \"%s%s.%d\"; transcript_id \"%s%s.%d.%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\"; sample_cnt \"%d\";\n",
		        chrom, t.exons[0].first, t.exons.back().second, t.strand, prefix, chrom, t.gene_id, prefix, chrom, t.gene_id, t.transcript_id,
		        k.fpkm, k.tpm, k.cov, k.support);
		for (size_t e = 0; e < t.exons.size(); ++e)
			fprintf(out, "%s\tPsiCLASS\texon\t%d\t%d\t1000\t%c\t.\tgene_id \"%s%s.%d\"; transcript_id \"%s%s.%d.%d\"; exon_number \"%d\"; FPKM \"%.6lf\"; TPM \"%.6lf\"; cov \"%.6lf\"; sample_cnt \"%d\";\n",
			        chrom, t.exons[e].first, t.exons[e].second, t.strand, prefix, chrom, t.gene_id, prefix, chrom, t.gene_id, t.transcript_id,
			        (int)e + 1, k.fpkm, k.tpm, k.cov, k.support);
	}
}

// The attribute parser of Vote.cpp:31-58, restated: the value of `field "<value>";` is the whitespace-delimited token that
// starts two characters after the first blank following the field name, cut at its last double quote.
static std::string gtf_value(const std::string &line, const char *field)
{
	size_t at = line.find(field);
	if (at == std::string::npos) return std::string();
	size_t blank = line.find(' ', at);
	if (blank == std::string::npos || blank + 2 > line.size()) return std::string();
	size_t from = blank + 2, to = from;
	while (to < line.size() && !isspace((unsigned char)line[to])) ++to;
	std::string tok = line.substr(from, to - from);
	size_t quote = tok.rfind('"');
	if (quote != std::string::npos) tok.erase(quote);
	return tok;
}
// the decimal number an id such as "chr2.17.3" ends with (Vote.cpp:47-58)
static int trailing_number(const std::string &id)
{
	size_t b = id.size();
	while (b > 0 && id[b - 1] >= '0' && id[b - 1] <= '9') --b;
	long long v = 0;
	for (size_t k = b; k < id.size(); ++k) v = v * 10 + (id[k] - '0');
	return (int)v;
}

bool vote_read_gtfs(const std::string &listPath, std::vector<VoteTx> &txs, int &sampleCnt, std::string &err)
{
	FILE *fpList = fopen(listPath.c_str(), "r");
	if (!fpList) {
		err = "cannot open " + listPath;
		return false;
	}
	char buffer[4096], line[10000], chrom[50], tool[20], type[40], strand[3];
	int start, end;
	sampleCnt = 0;
	VoteTx cur;
	bool have = false;
	while (fgets(buffer, sizeof(buffer), fpList) != NULL) {
		size_t len = strlen(buffer);
		if (len && buffer[len - 1] == '\n') buffer[--len] = '\0';
		FILE *fp = fopen(buffer, "r");
		if (!fp) {
			err = std::string("cannot open ") + buffer;
			fclose(fpList);
			return false;
		}
		auto flush = [&]() {
			if (have && !cur.exons.empty()) {
				cur.sample = sampleCnt;
				txs.push_back(cur);
			}
			cur.exons.clear();
			have = false;
		};
		while (fgets(line, sizeof(line), fp) != NULL) {
			if (line[0] == '#') continue;
			sscanf(line, "%s %s %s %d %d %s %s", chrom, tool, type, &start, &end, buffer, strand);
			if (strcmp(type, "exon")) {
				flush();
				continue;
			}
			const std::string text = line;
			cur.chrom = chrom;
			cur.gene_id = trailing_number(gtf_value(text, "gene_id"));
			cur.transcript_id = trailing_number(gtf_value(text, "transcript_id"));
			cur.fpkm = atof(gtf_value(text, "FPKM").c_str());
			cur.tpm = atof(gtf_value(text, "TPM").c_str());
			cur.cov = atof(gtf_value(text, "cov").c_str());
			cur.strand = strand[0];
			cur.exons.push_back({start, end});
			have = true;
		}
		flush();
		++sampleCnt;
		fclose(fp);
	}
	fclose(fpList);
	return true;
}

} // namespace psc
