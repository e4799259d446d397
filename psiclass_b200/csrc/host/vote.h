// Stage 4 of PsiCLASS, `vote-transcripts` (/root/reference/Vote.cpp:73-432): the per-sample transcripts are grouped by
// identical intron chain (MultiThreadOutputTranscript::CompSortTranscripts, TranscriptDecider.hpp:114-143), their
// FPKM / TPM / cov averaged, and a group is kept when enough samples support it or its summed coverage is high
// enough.  Two front ends share vote_core(): the drop-in `vote-transcripts` binary, which parses the stage-3 GTFs like
// the reference does, and `classes --vote <file>`, which feeds the transcripts it has just written (fused: no GTF
// print + parse between stage 3 and stage 4).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace psc {

struct VoteTx {
	int sample = 0, gene_id = 0, transcript_id = 0;
	std::string chrom;
	char strand = '.';
	std::vector<std::pair<int32_t, int32_t>> exons;
	double fpkm = 0, tpm = 0, cov = 0; // as printed in the GTF (six decimals)
};

// the value a "%.6lf" field parses back to (Vote.cpp:184-197 reads the numbers from the text)
double gtf_rounded(double x);

// `txs` in the order Vote.cpp reads them: sample by sample, each in the order of its GTF file
void vote_core(std::vector<VoteTx> &txs, int sample_cnt, double min_avg_depth, FILE *out);

// parses the GTFs listed in `list_path` the way Vote.cpp:130-247 does
bool vote_read_gtfs(const std::string &list_path, std::vector<VoteTx> &txs, int &sample_cnt, std::string &err);

} // namespace psc
