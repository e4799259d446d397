// ref_dump -- TEST INFRASTRUCTURE ONLY (oracle/): a small driver of our own that links the
// UNMODIFIED reference objects (Constraints.o, TranscriptDecider.o, SubexonGraph.o built by
// oracle/Makefile straight from /root/reference) and runs the same single-thread stage-3 loop
// the reference's classes.cpp:285-317 runs, while dumping the intermediate tables the parity
// tests pin: library stats (alignments.hpp:489-571), component list (SubexonGraph.cpp:43-112),
// per-(component,sample) constraint and mate-pair tables (Constraints.cpp:240-438; the
// reference's own commented-out dump is at Constraints.cpp:426-435) and totalReadCnt (Q8).
// The GTFs it writes are the reference's own (TranscriptDecider.hpp:230-283).
//
//   ref_dump -s subexon_combined.out --lb bamlist -o prefix [-c f] [-f f] [-d f]
//            [--maxDpConstraintSize n] [--primaryParalog] > dump.txt
//
// Nothing under psiclass_b200/ includes, links or executes this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "alignments.hpp"
#include "SubexonGraph.hpp"
#include "SubexonCorrelation.hpp"
#include "Constraints.hpp"
#include "TranscriptDecider.hpp"

static void dumpConstraints(FILE *fp, int comp, int sample, int seCnt, Constraints &cs)
{
	int C = (int)cs.constraints.size();
	int M = (int)cs.matePairs.size();
	fprintf(fp, "S %d %d %d %d\n", comp, sample, C, M);
	for (int k = 0; k < C; ++k) {
		struct _constraint &c = cs.constraints[k];
		fprintf(fp, "c %d %d %d %d %d %d %.17g %.17g", k, c.first, c.last, c.support,
		        c.uniqSupport, c.maxReadLen, c.weight, c.normAbund);
		std::vector<int> ones;
		c.vector.GetOnesIndices(ones);
		fprintf(fp, " %d", (int)ones.size());
		for (size_t t = 0; t < ones.size(); ++t)
			fprintf(fp, " %d", ones[t]);
		fprintf(fp, "\n");
	}
	for (int k = 0; k < M; ++k) {
		struct _matePairConstraint &m = cs.matePairs[k];
		fprintf(fp, "m %d %d %d %d %d %d %.17g\n", k, m.i, m.j, m.support, m.uniqSupport,
		        m.effectiveCount, m.normAbund);
	}
}

int main(int argc, char *argv[])
{
	const char *subexonPath = NULL, *bamList = NULL;
	char prefix[1024] = "";
	double fpkmFraction = 0.05, classifierThreshold = 0.05, minDepth = 2.5;
	int maxDp = 7;
	bool primaryParalog = false;
	bool quietTables = false;
	for (int a = 1; a < argc; ++a) {
		std::string o(argv[a]);
		if (o == "-s") subexonPath = argv[++a];
		else if (o == "--lb") bamList = argv[++a];
		else if (o == "-o") strcpy(prefix, argv[++a]);
		else if (o == "-c") classifierThreshold = atof(argv[++a]);
		else if (o == "-f") fpkmFraction = atof(argv[++a]);
		else if (o == "-d") minDepth = atof(argv[++a]);
		else if (o == "--maxDpConstraintSize") maxDp = atoi(argv[++a]);
		else if (o == "--primaryParalog") primaryParalog = true;
		else if (o == "--no-tables") quietTables = true;
		else { fprintf(stderr, "ref_dump: unknown option %s\n", argv[a]); return 1; }
	}
	if (!subexonPath || !bamList) { fprintf(stderr, "ref_dump: need -s and --lb\n"); return 1; }

	std::vector<Alignments> bams;
	{
		FILE *fp = fopen(bamList, "r");
		if (!fp) { fprintf(stderr, "ref_dump: cannot open %s\n", bamList); return 1; }
		char line[1024];
		while (fgets(line, sizeof(line), fp)) {
			size_t n = strlen(line);
			while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
			if (n == 0) continue;
			Alignments a;
			a.Open(line);
			bams.push_back(a);
		}
		fclose(fp);
	}
	int S = (int)bams.size();
	for (int s = 0; s < S; ++s) {
		bams[s].GetGeneralInfo(true);
		bams[s].Rewind();
		printf("L %d %d %d %d %d\n", s, bams[s].readLen, bams[s].fragLen, bams[s].fragStdev,
		       bams[s].matePaired ? 1 : 0);
	}

	FILE *fpSubexon = fopen(subexonPath, "r");
	if (!fpSubexon) { fprintf(stderr, "ref_dump: cannot open %s\n", subexonPath); return 1; }
	SubexonGraph graph(classifierThreshold, bams[0], fpSubexon);
	graph.ComputeGeneIntervals();

	std::vector<Constraints> sampleConstraints;
	for (int s = 0; s < S; ++s) {
		Constraints c(&bams[s]);
		c.SetHasMateReadIdSuffix(false);
		c.SetUsePrimaryAsUnique(primaryParalog);
		sampleConstraints.push_back(c);
	}
	SubexonCorrelation correlation;
	MultiThreadOutputTranscript out(S, bams[0]);
	out.SetOutputFPs(prefix);
	TranscriptDecider decider(fpkmFraction, classifierThreshold, minDepth, S, bams[0]);
	decider.SetMultiThreadOutputHandler(&out);
	decider.SetNumThreads(1);
	decider.SetMaxDpConstraintSize(maxDp);

	int nComp = (int)graph.geneIntervals.size();
	printf("G %d %d\n", nComp, (int)graph.subexons.size());
	for (int g = 0; g < nComp; ++g) {
		struct _geneInterval gi = graph.geneIntervals[g];
		int n = gi.endIdx - gi.startIdx + 1;
		struct _subexon *se = new struct _subexon[n];
		graph.ExtractSubexons(gi.startIdx, gi.endIdx, se);
		printf("C %d %d %d %d %d %d %d\n", g, n, se[0].chrId, gi.start, gi.end, se[0].lcCnt, se[0].rcCnt);
		if (!quietTables) {
			for (int k = 0; k < n; ++k) {
				printf("e %d %d %d %d %d %d %d %d %.17g %.17g %d", k, se[k].start, se[k].end,
				       se[k].leftType, se[k].rightType, se[k].leftStrand, se[k].rightStrand,
				       se[k].geneId, se[k].leftClassifier, se[k].rightClassifier, se[k].prevCnt);
				for (int t = 0; t < se[k].prevCnt; ++t) printf(" %d", se[k].prev[t]);
				printf(" %d", se[k].nextCnt);
				for (int t = 0; t < se[k].nextCnt; ++t) printf(" %d", se[k].next[t]);
				printf("\n");
			}
		}
		correlation.ComputeCorrelation(se, n, bams[0]);
		for (int s = 0; s < S; ++s) {
			sampleConstraints[s].BuildConstraints(se, n, gi.start, gi.end);
			if (!quietTables)
				dumpConstraints(stdout, g, s, n, sampleConstraints[s]);
		}
		decider.Solve(se, n, sampleConstraints, correlation);
		for (int k = 0; k < n; ++k) { delete[] se[k].prev; delete[] se[k].next; }
		delete[] se;
	}
	for (int s = 0; s < S; ++s)
		printf("T %d %d\n", s, bams[s].totalReadCnt);

	out.OutputCommandInfo(argc, argv);
	for (int s = 0; s < S; ++s) {
		char name[1024];
		bams[s].GetFileName(name);
		char *base = strrchr(name, '/');
		out.OutputCommentToSampleGTF(s, base ? base + 1 : name);
	}
	out.ComputeFPKMTPM(bams);
	out.Flush();
	fflush(stdout);
	fclose(fpSubexon);
	return 0;
}
