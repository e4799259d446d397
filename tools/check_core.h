// check_core.h -- TEST TOOL: engine-vs-oracle comparison shared by psc_check and the pytest testkit
#pragma once
#include <vector>
#include "../include/psc.h"
#include "../psiclass_b200/csrc/host/bam_io.h"
#include "../psiclass_b200/csrc/host/subexon_graph.h"

struct EngineApi {
	int (*create)(psc_ctx **, int);
	void (*destroy)(psc_ctx *);
	const char *(*last_error)(const psc_ctx *);
	int (*set_params)(psc_ctx *, const psc_params *);
	int (*set_graph)(psc_ctx *, const psc_graph *);
	int (*set_reads)(psc_ctx *, int32_t, const psc_reads *);
	int (*build_constraints)(psc_ctx *);
	int (*constraint_counts)(psc_ctx *, int32_t *, int32_t *);
	int (*get_constraint_table)(psc_ctx *, int32_t, int32_t, psc_constraint_table *);
	int (*solve)(psc_ctx *);
	int (*transcript_counts)(psc_ctx *, int64_t *, int64_t *);
	int (*get_transcripts)(psc_ctx *, psc_transcripts *);
	int (*get_comp_stats)(psc_ctx *, psc_comp_stats *);
	int (*get_timing)(psc_ctx *, psc_timing *);
	int (*set_coop_width)(psc_ctx *, int32_t);
};
struct CheckOpts { int maxDp = 7, coop = 32, primaryParalog = 0, constraintsOnly = 0, verbose = 0, device = 0; };
struct CheckReport {
	int pass, n_comp, n_dp;
	long long tables_checked, tables_bad, stats_bad, tx_engine, tx_oracle, tx_bad, val_exact, val_close, val_bad, launches;
	double constraints_ms, solve_ms;
	long long n_downsampled, n_iter10, n_skipped; // oracle-side branch counters (orc::SolveStats)
};
int run_check(const psc::Graph &g, std::vector<psc::SampleReads> &samples, const EngineApi &api, const CheckOpts &opt, CheckReport &rep);
