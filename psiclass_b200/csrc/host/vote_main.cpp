// vote-transcripts -- drop-in for the reference's stage-4 executable (/root/reference/Vote.cpp): same argv
// (--lg <list of GTF files>, -d <min average depth>), vote on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vote.h"

static const char usage[] = "./transcript-vote [OPTIONS] > output.gtf:\n"
                            "Required:\n"
                            "\t--lg: path to the list of GTF files.\n"
                            "Optional:\n"
                            "\t-d FLOAT: threshold of average coverage depth across all the samples. (default: 1)\n";

int main(int argc, char *argv[])
{
	double minAvgDepth = 1.0;
	const char *list = NULL;
	for (int i = 1; i < argc; ++i) {
		if (!strcmp(argv[i], "--lg") && i + 1 < argc) list = argv[++i];
		else if (!strcmp(argv[i], "-d") && i + 1 < argc) minAvgDepth = atof(argv[++i]);
		else {
			printf("%s", usage);
			exit(1);
		}
	}
	if (!list) {
		printf("Must use --lg option to speicfy the list of GTF files.\n%s", usage);
		exit(1);
	}
	std::vector<psc::VoteTx> txs;
	int sampleCnt = 0;
	std::string err;
	if (!psc::vote_read_gtfs(list, txs, sampleCnt, err)) {
		fprintf(stderr, "%s\n", err.c_str());
		exit(1);
	}
	psc::vote_core(txs, sampleCnt, minAvgDepth, stdout);
	return 0;
}
