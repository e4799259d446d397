// psc_check -- TEST TOOL: runs the engine behind the C ABI (libpsc.so on a GPU box, or the CPU
// shim build under tests/shim/_build when debugging without a GPU) and the CPU oracle on the same
// dataset and compares every table (see check_core.cpp).
// usage: psc_check <dir with subexon_combined.out + bamlist> [--maxDp n] [--coop 1|32] [--primaryParalog]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "check_core.h"

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: psc_check dir [opts]\n"); return 2; }
	std::string dir = argv[1];
	CheckOpts opt;
	for (int a = 2; a < argc; ++a) {
		std::string o = argv[a];
		if (o == "--maxDp") opt.maxDp = atoi(argv[++a]);
		else if (o == "--coop") opt.coop = atoi(argv[++a]);
		else if (o == "--primaryParalog") opt.primaryParalog = 1;
		else if (o == "--constraints-only") opt.constraintsOnly = 1;
		else if (o == "-v") opt.verbose = 1;
	}
	std::vector<std::string> bams;
	{
		FILE *fp = fopen((dir + "/bamlist").c_str(), "r");
		if (!fp) { fprintf(stderr, "no bamlist\n"); return 2; }
		char line[1024];
		while (fgets(line, sizeof(line), fp)) {
			size_t n = strlen(line);
			while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
			if (n) bams.push_back(line);
		}
		fclose(fp);
	}
	std::vector<psc::SampleReads> samples(bams.size());
	std::string err;
	for (size_t s = 0; s < bams.size(); ++s)
		if (!psc::decode_bam(bams[s], 4, samples[s], err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
	psc::Graph g;
	if (!psc::load_subexon_graph(dir + "/subexon_combined.out", samples[0].chr_names, 0.05, g, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
	EngineApi api = {psc_create, psc_destroy, psc_last_error, psc_set_params, psc_set_graph, psc_set_reads, psc_build_constraints,
	                 psc_constraint_counts, psc_get_constraint_table, psc_solve, psc_transcript_counts, psc_get_transcripts,
	                 psc_get_comp_stats, psc_get_timing, psc_set_coop_width};
	CheckReport rep;
	int rc = run_check(g, samples, api, opt, rep);
	printf(rep.pass ? "RESULT: PASS\n" : "RESULT: FAIL\n");
	return rc;
}
