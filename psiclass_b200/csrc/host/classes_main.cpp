// classes -- drop-in replacement for the reference's stage-3 executable (classes.cpp:97-508):
// same argv surface, same inputs (combined subexon file + coordinate-sorted BAMs), same outputs
// (<prefix>_sample_<i>.gtf consumed by vote-transcripts; per-component log lines on stdout).
// All per-component work runs on the GPU(s) behind the C ABI of include/psc.h; the pthread
// reader/worker driver of the reference becomes a batch scheduler that shards the components
// over the visible devices (environment variable PSC_GPUS, default 1), with no collective on the
// path: each device gets a contiguous block of components with all samples' reads for them.
#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/psc.h"
#include "bam_io.h"
#include "gtf_out.h"
#include "subexon_graph.h"

static const char usage[] =
    "./classes [OPTIONS]:\n"
    "Required:\n"
    "\t-s STRING: path to the subexon file.\n"
    "\t-b STRING: path to the BAM file.\n"
    "\t\tor\n"
    "\t--lb STRING: path to the file of the list of BAM files.\n"
    "Optional:\n"
    "\t-p INT: number of host threads used to decode the BAM files. (default: 1)\n"
    "\t-o STRING: the prefix of the output file. (default: not used)\n"
    "\t-c FLOAT: only use the subexons with classifier score <= than the given number. (default: 0.05)\n"
    "\t-f FLOAT: filter the transcript from the gene if its abundance is lower than the given number percent of the most abundant one. (default: 0.05)\n"
    "\t-d FLOAT: filter the transcript whose average read depth is less than the given number. (default: 2.5)\n"
    "\t--ls STRING: path to the file of the list of single-sample subexon files. (default: not used)\n"
    "\t--stranded STRING: un/rf/fr for library unstranded/fr-firstand/fr-secondstrand (default: not used)\n"
    "\t--hasMateIdSuffix: the read id has suffix such as .1, .2 for a mate pair. (default: false)\n"
    "\t--maxDpConstraintSize: the maximum number of subexons a constraint can cover in dynamic programming. (default: 7; -1 for inf)\n"
    "\t--primaryParalog: use primary alignment to retain paralog genes instead of unique alignments. (default: not used)\n"
    "\t--vote STRING: also write the stage-4 vote (what vote-transcripts would make of the GTFs) to this file. (default: not used)\n"
    "\t--vote-depth FLOAT: vote-transcripts' -d. (default: 1)\n";

static const char *short_options = "s:b:f:o:d:p:c:h";
static struct option long_options[] = {{"ls", required_argument, 0, 10000},
                                       {"lb", required_argument, 0, 10001},
                                       {"hasMateIdSuffix", no_argument, 0, 10002},
                                       {"primaryParalog", no_argument, 0, 10003},
                                       {"maxDpConstraintSize", required_argument, 0, 10004},
                                       {"stranded", required_argument, 0, 10005},
                                       {"vote", required_argument, 0, 10006},
                                       {"vote-depth", required_argument, 0, 10007},
                                       {(char *)0, 0, 0, 0}};

struct Shard {
	int c0 = 0, c1 = 0;
	std::vector<int32_t> comp_off, next_off, prev_off, next_idx, prev_idx;
	std::vector<std::vector<int64_t>> chr_off; // per sample
	std::vector<int64_t> lo, hi;               // per sample read range
	std::vector<psc::OutTx> out;
	std::vector<int32_t> at_cnt, use_dp, ncon0, npair0;
	std::string err;
	double upload_s = 0, constraints_s = 0, solve_s = 0, download_s = 0; // PSC_PERF_JSON
	long long n_reads = 0;
};

static double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Device contexts are created while the BAM files are still being decoded (context creation takes about a second).
struct EarlyCtx {
	psc_ctx *ctx = nullptr;
	int rc = 0;
	std::thread th;
};

static void run_shard(int device, const psc::Graph &g, const std::vector<psc::SampleReads> &samples, const psc_params &prm,
                      Shard &sh, EarlyCtx *early)
{
	int S = (int)samples.size();
	int nc = sh.c1 - sh.c0;
	if (nc <= 0) return;
	int se0 = g.comp_off[(size_t)sh.c0], se1 = g.comp_off[(size_t)sh.c1];
	int ne0 = g.next_off[(size_t)se0], pe0 = g.prev_off[(size_t)se0];
	sh.comp_off.resize((size_t)nc + 1);
	for (int c = 0; c <= nc; ++c) sh.comp_off[(size_t)c] = g.comp_off[(size_t)(sh.c0 + c)] - se0;
	sh.next_off.resize((size_t)(se1 - se0) + 1);
	sh.prev_off.resize((size_t)(se1 - se0) + 1);
	for (int i = se0; i <= se1; ++i) {
		sh.next_off[(size_t)(i - se0)] = g.next_off[(size_t)i] - ne0;
		sh.prev_off[(size_t)(i - se0)] = g.prev_off[(size_t)i] - pe0;
	}
	psc_ctx *ctx = nullptr;
	int rc;
	if (early) {
		if (early->th.joinable()) early->th.join();
		ctx = early->ctx;
		rc = early->rc;
		early->ctx = nullptr;
	} else
		rc = psc_create(&ctx, device);
	if (rc) {
		sh.err = "psc_create failed (code " + std::to_string(rc) + "): a CUDA device is required";
		return;
	}
#define E(call)                                         \
	do {                                                 \
		if ((call) != 0) {                                \
			sh.err = std::string(#call) + ": " + psc_last_error(ctx); \
			psc_destroy(ctx);                              \
			return;                                        \
		}                                                 \
	} while (0)
	E(psc_set_params(ctx, &prm));
	psc_graph pg;
	pg.n_comp = nc;
	pg.n_se = se1 - se0;
	pg.comp_off = sh.comp_off.data();
	pg.comp_chr = g.comp_chr.data() + sh.c0;
	pg.comp_base_gene = g.comp_base_gene.data() + sh.c0;
	pg.comp_used_gene = g.comp_used_gene.data() + sh.c0;
	pg.se_start = g.se_start.data() + se0;
	pg.se_end = g.se_end.data() + se0;
	pg.se_left_type = g.se_left_type.data() + se0;
	pg.se_right_type = g.se_right_type.data() + se0;
	pg.se_left_strand = g.se_left_strand.data() + se0;
	pg.se_right_strand = g.se_right_strand.data() + se0;
	pg.se_left_cls = g.se_left_cls.data() + se0;
	pg.se_right_cls = g.se_right_cls.data() + se0;
	pg.se_gene_id = g.se_gene_id.data() + se0;
	pg.next_off = sh.next_off.data();
	pg.next_idx = g.next_idx.data() + ne0;
	pg.prev_off = sh.prev_off.data();
	pg.prev_idx = g.prev_idx.data() + pe0;
	E(psc_set_graph(ctx, &pg));
	double t0 = now_s();
	// reads of the shard: [first read the first component examines, last read the last component examines)
	sh.chr_off.resize((size_t)S);
	for (int s = 0; s < S; ++s) {
		const psc::SampleReads &sr = samples[(size_t)s];
		int64_t lo = sh.lo[(size_t)s], hi = sh.hi[(size_t)s];
		int nchr = (int)sr.chr_names.size();
		std::vector<int64_t> &co = sh.chr_off[(size_t)s];
		co.resize((size_t)nchr + 1);
		for (int k = 0; k <= nchr; ++k) co[(size_t)k] = std::min(std::max(sr.chr_off[(size_t)k], lo), hi) - lo;
		psc_reads r;
		r.n_reads = hi - lo;
		r.n_chr = nchr;
		r.chr_off = co.data();
		r.pos = sr.pos.data() + lo;
		r.mate_pos = sr.mate_pos.data() + lo;
		r.flags = sr.flags.data() + lo;
		r.nh = sr.nh.data() + lo;
		r.name_key = sr.name_key.data() + lo;
		// seg_off is used relative to seg_a/seg_b; keep the absolute offsets and the full arrays
		r.seg_off = sr.seg_off.data() + lo;
		r.seg_a = sr.seg_a.data();
		r.seg_b = sr.seg_b.data();
		r.n_segs = (int64_t)sr.seg_a.size();
		r.read_len = sr.read_len;
		r.frag_stdev = sr.frag_stdev;
		E(psc_set_reads(ctx, s, &r));
		sh.n_reads += r.n_reads;
	}
	double t1 = now_s();
	E(psc_build_constraints(ctx));
	double t2 = now_s();
	E(psc_solve(ctx));
	double t3 = now_s();
	sh.upload_s = t1 - t0;
	sh.constraints_s = t2 - t1;
	sh.solve_s = t3 - t2;
	std::vector<int32_t> nC((size_t)S * nc), nM((size_t)S * nc);
	E(psc_constraint_counts(ctx, nC.data(), nM.data()));
	sh.ncon0.assign(nC.begin(), nC.begin() + nc);
	sh.npair0.assign(nM.begin(), nM.begin() + nc);
	sh.at_cnt.resize((size_t)nc);
	sh.use_dp.resize((size_t)nc);
	std::vector<int32_t> ncd((size_t)nc);
	psc_comp_stats st;
	st.at_cnt = sh.at_cnt.data();
	st.use_dp = sh.use_dp.data();
	st.n_candidates = ncd.data();
	E(psc_get_comp_stats(ctx, &st));
	int64_t nt = 0, ne = 0;
	E(psc_transcript_counts(ctx, &nt, &ne));
	std::vector<int32_t> tc((size_t)nt + 1), ts((size_t)nt + 1), tchr((size_t)nt + 1), tg((size_t)nt + 1), es((size_t)ne + 1), ee((size_t)ne + 1);
	std::vector<int8_t> tstr((size_t)nt + 1);
	std::vector<int64_t> eo((size_t)nt + 2);
	std::vector<double> tf((size_t)nt + 1), tcov((size_t)nt + 1);
	psc_transcripts pt;
	pt.comp = tc.data(); pt.sample = ts.data(); pt.chr = tchr.data(); pt.gene_id = tg.data(); pt.strand = tstr.data();
	pt.exon_off = eo.data(); pt.exon_start = es.data(); pt.exon_end = ee.data(); pt.fpkm = tf.data(); pt.cov = tcov.data();
	E(psc_get_transcripts(ctx, &pt));
	sh.out.resize((size_t)nt);
	for (int64_t k = 0; k < nt; ++k) {
		psc::OutTx &o = sh.out[(size_t)k];
		o.comp = tc[(size_t)k] + sh.c0;
		o.sample = ts[(size_t)k];
		o.chr = tchr[(size_t)k];
		o.gene_id = tg[(size_t)k];
		o.strand = (char)tstr[(size_t)k];
		o.fpkm = tf[(size_t)k];
		o.cov = tcov[(size_t)k];
		o.tpm = 0;
		for (int64_t e = eo[(size_t)k]; e < eo[(size_t)k + 1]; ++e) o.exons.push_back({es[(size_t)e], ee[(size_t)e]});
	}
	sh.download_s = now_s() - t3;
	// the context is left to process exit: freeing its buffers one by one costs most of a second per device
	(void)ctx;
#undef E
}

int main(int argc, char *argv[])
{
	if (argc <= 1) {
		printf("%s", usage);
		return 0;
	}
	std::string subexonPath, prefix;
	std::vector<std::string> bamPaths;
	double fpkmFraction = 0.05, classifierThreshold = 0.05, minDepth = 2.5;
	int numThreads = 1, maxDp = 7;
	bool primaryParalog = false;
	std::string votePath;
	double voteDepth = 1.0;
	int c, option_index = 0;
	while ((c = getopt_long(argc, argv, short_options, long_options, &option_index)) != -1) {
		if (c == 's') subexonPath = optarg;
		else if (c == 'b') bamPaths.push_back(optarg);
		else if (c == 'c') classifierThreshold = atof(optarg);
		else if (c == 'f') fpkmFraction = atof(optarg);
		else if (c == 'o') prefix = optarg;
		else if (c == 'd') minDepth = atof(optarg);
		else if (c == 'p') numThreads = atoi(optarg);
		else if (c == 10000) { /* --ls: subexon correlation is inert on the psiclass path */ }
		else if (c == 10001) {
			FILE *fp = fopen(optarg, "r");
			if (!fp) {
				printf("Cannot open %s.\n", optarg);
				exit(1);
			}
			char buffer[1024];
			while (fgets(buffer, sizeof(buffer), fp) != NULL) {
				size_t len = strlen(buffer);
				while (len && (buffer[len - 1] == '\n' || buffer[len - 1] == '\r')) buffer[--len] = 0;
				if (len) bamPaths.push_back(buffer);
			}
			fclose(fp);
		} else if (c == 10002) { /* --hasMateIdSuffix: irrelevant, the suffix rule is always on (Q2) */ }
		else if (c == 10003) primaryParalog = true;
		else if (c == 10004) maxDp = atoi(optarg);
		else if (c == 10005) { /* --stranded only changes stages 0-2 (Q13) */ }
		else if (c == 10006) votePath = optarg;
		else if (c == 10007) voteDepth = atof(optarg);
		else {
			printf("%s", usage);
			exit(1);
		}
	}
	if (subexonPath.empty() || !fopen(subexonPath.c_str(), "r")) {
		printf("Cannot find combined subexon file.\n");
		exit(1);
	}
	if (bamPaths.empty()) {
		printf("Must use -b option to specify BAM files.\n");
		exit(1);
	}
	int S = (int)bamPaths.size();
	if (numThreads < 1) numThreads = 1;
	const double tStart = now_s();
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); // psc_solve drives several streams at once (before CUDA initialises)
	int nGpu = 1;
	if (const char *e = getenv("PSC_GPUS")) nGpu = std::max(1, atoi(e));
	std::vector<EarlyCtx> early((size_t)nGpu);
	for (int d = 0; d < nGpu; ++d) early[(size_t)d].th = std::thread([&early, d]() { early[(size_t)d].rc = psc_create(&early[(size_t)d].ctx, d); });
	// ---- the subexon graph is parsed while the BAMs decode (it only needs the chromosome names of BAM #0's header) ----
	psc::Graph g;
	std::string graphErr;
	bool graphOk = false;
	std::thread graphThread([&]() {
		std::vector<std::string> names;
		std::vector<int32_t> lens;
		if (!psc::read_bam_header(bamPaths[0], names, lens, graphErr)) return;
		// PSC_GRAPH_CACHE=1: the parsed graph is kept as <subexon file>.pscgraph and mapped back on the next run
		const char *gc = getenv("PSC_GRAPH_CACHE");
		const bool useCache = gc && atoi(gc) != 0;
		if (useCache && psc::load_graph_cache(subexonPath + ".pscgraph", subexonPath, names, classifierThreshold, g)) {
			graphOk = true;
			return;
		}
		graphOk = psc::load_subexon_graph(subexonPath, names, classifierThreshold, g, graphErr);
		if (graphOk && useCache) psc::save_graph_cache(subexonPath + ".pscgraph", subexonPath, classifierThreshold, g);
	});
	// ---- one decode pass per BAM (library statistics + alignment records), samples in parallel ----
	std::vector<psc::SampleReads> samples((size_t)S);
	std::vector<std::string> errs((size_t)S);
	{
		int outer = std::min(numThreads, S), inner = std::max(1, (numThreads + outer - 1) / std::max(1, outer)); // inflate workers per sample (the parser thread comes on top)
		std::vector<std::thread> pool;
		for (int t = 0; t < outer; ++t)
			pool.emplace_back([&, t]() {
				for (int s = t; s < S; s += outer)
					if (!psc::decode_bam(bamPaths[(size_t)s], inner, samples[(size_t)s], errs[(size_t)s]) && errs[(size_t)s].empty())
						errs[(size_t)s] = "decode failed";
			});
		for (auto &t : pool) t.join();
	}
	for (int s = 0; s < S; ++s)
		if (!errs[(size_t)s].empty()) {
			fprintf(stderr, "%s\n", errs[(size_t)s].c_str());
			exit(1);
		}
	const double tDecoded = now_s();
	graphThread.join();
	std::string err;
	if (!graphOk) {
		printf("%s\n", graphErr.c_str());
		exit(1);
	}
	psc_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.fpkm_fraction = fpkmFraction;
	prm.classifier_threshold = classifierThreshold;
	prm.min_read_depth = minDepth;
	prm.max_dp_constraint_size = maxDp;
	prm.primary_as_unique = primaryParalog;
	prm.n_samples = S;
	prm.read_len = samples[0].read_len;
	prm.frag_len = samples[0].frag_len;
	prm.frag_stdev = samples[0].frag_stdev;
	prm.mate_paired = samples[0].mate_paired;

	const double tGraph = now_s();
	// ---- shard the components over the devices, balanced by an estimate of the device work ----
	const int nGpuAsked = nGpu;
	nGpu = std::min(nGpu, std::max(1, g.n_comp));
	// per component and sample: the first read it examines (Constraints.cpp:272-283 in closed form)
	std::vector<int> maxEnd((size_t)g.n_comp);
	for (int c2 = 0; c2 < g.n_comp; ++c2) {
		int m = INT_MIN;
		for (int i = g.comp_off[(size_t)c2]; i < g.comp_off[(size_t)c2 + 1]; ++i) m = std::max(m, g.se_end[(size_t)i]);
		maxEnd[(size_t)c2] = m;
	}
	auto first_read = [&](int s, int comp) -> int64_t { // lo of the component's group
		const psc::SampleReads &sr = samples[(size_t)s];
		int chr = g.comp_chr[(size_t)comp];
		int64_t a = sr.chr_off[(size_t)chr], b = sr.chr_off[(size_t)chr + 1];
		if (comp > 0 && g.comp_chr[(size_t)comp - 1] == chr)
			return std::upper_bound(sr.pos.begin() + a, sr.pos.begin() + b, maxEnd[(size_t)comp - 1]) - sr.pos.begin();
		return a;
	};
	auto last_read = [&](int s, int comp) -> int64_t {
		const psc::SampleReads &sr = samples[(size_t)s];
		int chr = g.comp_chr[(size_t)comp];
		int64_t a = sr.chr_off[(size_t)chr], b = sr.chr_off[(size_t)chr + 1];
		return std::upper_bound(sr.pos.begin() + a, sr.pos.begin() + b, maxEnd[(size_t)comp]) - sr.pos.begin();
	};
	std::vector<Shard> shards((size_t)nGpu);
	{
		std::vector<long long> cum((size_t)g.n_comp + 1, 0);
		for (int c2 = 0; c2 < g.n_comp; ++c2) {
			long long r = 16;
			for (int s = 0; s < S; ++s) r += last_read(s, c2) - first_read(s, c2);
			// constraint building is linear in the reads; candidate search and set cover grow with the number of
			// subexons as well (the DP components, n >~ 20, carry most of the solve time)
			long long n = g.comp_off[(size_t)c2 + 1] - g.comp_off[(size_t)c2];
			cum[(size_t)c2 + 1] = cum[(size_t)c2] + r * (8 + n) / 8;
		}
		int c0 = 0;
		for (int d = 0; d < nGpu; ++d) {
			long long target = cum[(size_t)g.n_comp] * (d + 1) / nGpu;
			int c1 = d == nGpu - 1 ? g.n_comp : (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
			c1 = std::max(c1, c0);
			c1 = std::min(c1, g.n_comp);
			shards[(size_t)d].c0 = c0;
			shards[(size_t)d].c1 = c1;
			shards[(size_t)d].lo.resize((size_t)S);
			shards[(size_t)d].hi.resize((size_t)S);
			for (int s = 0; s < S; ++s) {
				shards[(size_t)d].lo[(size_t)s] = c1 > c0 ? first_read(s, c0) : 0;
				shards[(size_t)d].hi[(size_t)s] = c1 > c0 ? last_read(s, c1 - 1) : 0;
			}
			c0 = c1;
		}
	}
	{
		std::vector<std::thread> pool;
		for (int d = 0; d < nGpu; ++d) pool.emplace_back([&, d]() { run_shard(d, g, samples, prm, shards[(size_t)d], &early[(size_t)d]); });
		for (auto &t : pool) t.join();
	}
	for (int d = 0; d < nGpuAsked; ++d) { // contexts of devices that got no shard
		if (early[(size_t)d].th.joinable()) early[(size_t)d].th.join();
		if (early[(size_t)d].ctx) psc_destroy(early[(size_t)d].ctx);
	}
	for (auto &sh : shards)
		if (!sh.err.empty()) {
			printf("%s\n", sh.err.c_str());
			exit(1);
		}
	const double tDevice = now_s();
	// ---- log (classes.cpp:299-302, TranscriptDecider.cpp:3111) and output ----
	std::vector<psc::OutTx> queue;
	for (auto &sh : shards) {
		for (int k = 0; k < sh.c1 - sh.c0; ++k) {
			int comp = sh.c0 + k;
			int n = g.comp_off[(size_t)comp + 1] - g.comp_off[(size_t)comp];
			printf("%d: %d %s %d %d\n", comp, n, g.chr_names[(size_t)g.comp_chr[(size_t)comp]].c_str(), g.comp_start[(size_t)comp] + 1,
			       g.comp_end[(size_t)comp] + 1);
			printf("%d: atCnt=%d seCnt=%d %d %d %d\n", g.se_start[(size_t)g.comp_off[(size_t)comp]] + 1, sh.at_cnt[(size_t)k], n,
			       sh.use_dp[(size_t)k], sh.ncon0[(size_t)k], sh.npair0[(size_t)k]);
		}
		for (auto &t : sh.out) queue.push_back(std::move(t));
	}
	// Q8: totalReadCnt = #primary records read when the stream stopped after the last component
	std::vector<double> totalReadCnt((size_t)S, 0.0);
	for (int s = 0; s < S; ++s) {
		const psc::SampleReads &sr = samples[(size_t)s];
		if (g.n_comp == 0) continue;
		int64_t x = last_read(s, g.n_comp - 1);
		totalReadCnt[(size_t)s] = x < sr.n_reads() ? (double)sr.prim_cum[(size_t)x] : (double)sr.total_primary;
	}
	std::vector<psc::VoteTx> voteTx;
	if (!psc::write_gtfs(queue, totalReadCnt, g.chr_names, prefix, argc, argv, bamPaths, err, votePath.empty() ? nullptr : &voteTx)) {
		printf("%s\n", err.c_str());
		exit(1);
	}
	const double tGtf = now_s();
	if (!votePath.empty()) { // stage 4 fused: Vote.cpp on the transcripts in memory
		FILE *fp = fopen(votePath.c_str(), "w");
		if (!fp) {
			printf("Cannot write %s.\n", votePath.c_str());
			exit(1);
		}
		psc::vote_core(voteTx, S, voteDepth, fp);
		fclose(fp);
	}
	if (const char *pj = getenv("PSC_PERF_JSON")) { // wall-clock of the phases, for bench.py (path of the file to write)
		if (FILE *fp = fopen(pj, "w")) {
			long long bamBytes = 0, nReads = 0;
			for (auto &b : bamPaths)
				if (FILE *f = fopen(b.c_str(), "rb")) {
					fseek(f, 0, SEEK_END);
					bamBytes += ftell(f);
					fclose(f);
				}
			for (auto &sr : samples) nReads += sr.n_reads();
			const double tEnd = now_s();
			fprintf(fp, "{\"components\": %d, \"samples\": %d, \"gpus\": %d, \"host_threads\": %d, \"reads\": %lld, \"bam_bytes\": %lld, "
			            "\"decode_s\": %.4f, \"decode_MBps\": %.1f, \"graph_s\": %.4f, \"device_s\": %.4f, \"gtf_s\": %.4f, \"vote_s\": %.4f, \"total_s\": %.4f, \"shards\": [",
			        g.n_comp, S, nGpu, numThreads, nReads, bamBytes, tDecoded - tStart, bamBytes / 1e6 / std::max(1e-9, tDecoded - tStart),
			        tGraph - tDecoded, tDevice - tGraph, tGtf - tDevice, tEnd - tGtf, tEnd - tStart);
			for (size_t d = 0; d < shards.size(); ++d)
				fprintf(fp, "%s{\"components\": %d, \"reads\": %lld, \"upload_s\": %.4f, \"constraints_s\": %.4f, \"solve_s\": %.4f, \"download_s\": %.4f}",
				        d ? ", " : "", shards[d].c1 - shards[d].c0, shards[d].n_reads, shards[d].upload_s, shards[d].constraints_s,
				        shards[d].solve_s, shards[d].download_s);
			fprintf(fp, "]}\n");
			fclose(fp);
		}
	}
	return 0;
}
