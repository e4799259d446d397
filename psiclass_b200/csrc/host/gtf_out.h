// GTF writer of stage 3: MultiThreadOutputTranscript::ComputeFPKMTPM / Flush
// (/root/reference/TranscriptDecider.hpp:190-283) and the '#' header lines
// (TranscriptDecider.hpp:210-228, classes.cpp:479-500).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vote.h"

namespace psc {

struct OutTx {
	int32_t comp, sample, chr, gene_id;
	char strand;
	std::vector<std::pair<int32_t, int32_t>> exons; // 1-based inclusive
	double fpkm, cov, tpm;
};

// `queue` must be in the reference's queue order (component-major, then sample, then pick order);
// total_read_cnt[s] is Alignments::totalReadCnt of sample s when the stream stopped (Q8).
bool write_gtfs(std::vector<OutTx> &queue, const std::vector<double> &total_read_cnt,
                const std::vector<std::string> &chr_names, const std::string &prefix, int argc, char **argv,
                const std::vector<std::string> &bam_paths, std::string &err, std::vector<VoteTx> *for_vote = nullptr);
// for_vote: when given, receives what vote-transcripts would parse back from the files just written (sample by
// sample, in file order, numbers rounded to the six printed decimals) -- the fused stage 3 -> stage 4 hand-over

} // namespace psc
