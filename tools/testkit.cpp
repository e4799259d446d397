// testkit -- TEST INFRASTRUCTURE for pytest and bench.py: synthetic workloads, file loading, the
// CPU oracle and the engine-vs-oracle comparison behind a small C API (loaded with ctypes).
// The engine under test is dlopen()ed from the path the caller gives (psiclass_b200/libpsc.so on a
// GPU box; the CPU shim only in the "not gpu" logic tests), so everything goes through the C ABI.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "check_core.h"
#include "synth.h"

struct tk_handle {
	synth::Workload w;
	bool synthetic = false;
	psc::Graph g;
	std::vector<psc::SampleReads> samples;
	std::string err;
};

struct tk_synth_cfg {
	uint64_t seed;
	int32_t n_chr, n_genes, n_samples;
	double mean_pairs, exon_mu, exon_sigma;
	int32_t exon_min, exon_max;
	double iso_lambda, p_skip, p_alt, p_ir, p_noise, p_overlap, p_paralog;
	int32_t name_style, stranded, dense, graph_noise, single_end, threads;
	double p_nh3, frag_sd;
};

static bool graph_from_text(tk_handle *h, const std::string &txt)
{
	char tmpl[] = "/tmp/psc_tk_XXXXXX";
	int fd = mkstemp(tmpl);
	if (fd < 0) { h->err = "mkstemp failed"; return false; }
	FILE *fp = fdopen(fd, "w");
	fwrite(txt.data(), 1, txt.size(), fp);
	fclose(fp);
	bool ok = psc::load_subexon_graph(tmpl, h->w.chr_names, 0.05, h->g, h->err);
	remove(tmpl);
	return ok;
}

extern "C" {

void tk_default_cfg(tk_synth_cfg *c)
{
	synth::Config d;
	c->seed = d.seed; c->n_chr = d.n_chr; c->n_genes = d.n_genes; c->n_samples = d.n_samples;
	c->mean_pairs = d.mean_pairs; c->exon_mu = d.exon_mu; c->exon_sigma = d.exon_sigma;
	c->exon_min = d.exon_min; c->exon_max = d.exon_max; c->iso_lambda = d.iso_lambda; c->p_skip = d.p_skip;
	c->p_alt = d.p_alt; c->p_ir = d.p_ir; c->p_noise = d.p_noise; c->p_overlap = d.p_overlap_gene; c->p_paralog = d.p_paralog;
	c->name_style = d.name_style; c->stranded = d.stranded; c->dense = d.dense; c->graph_noise = d.graph_noise;
	c->single_end = d.single_end; c->threads = 8;
	c->p_nh3 = d.p_nh3; c->frag_sd = d.frag_sd;
}

tk_handle *tk_synth(const tk_synth_cfg *c)
{
	tk_handle *h = new tk_handle;
	synth::Config d;
	d.seed = c->seed; d.n_chr = c->n_chr; d.n_genes = c->n_genes; d.n_samples = c->n_samples;
	d.mean_pairs = c->mean_pairs; d.exon_mu = c->exon_mu; d.exon_sigma = c->exon_sigma;
	d.exon_min = c->exon_min; d.exon_max = c->exon_max; d.iso_lambda = c->iso_lambda; d.p_skip = c->p_skip;
	d.p_alt = c->p_alt; d.p_ir = c->p_ir; d.p_noise = c->p_noise; d.p_overlap_gene = c->p_overlap; d.p_paralog = c->p_paralog;
	d.name_style = c->name_style; d.stranded = c->stranded; d.dense = c->dense; d.graph_noise = c->graph_noise;
	d.single_end = c->single_end;
	d.p_nh3 = c->p_nh3; d.frag_sd = c->frag_sd;
	synth::make_genes(d, h->w);
	h->synthetic = true;
	if (!graph_from_text(h, synth::subexon_text(h->w))) return h;
	h->samples.resize((size_t)d.n_samples);
	int T = c->threads > 0 ? c->threads : 1;
	std::vector<std::thread> pool;
	for (int t = 0; t < T; ++t)
		pool.emplace_back([&, t]() {
			for (int s = t; s < d.n_samples; s += T) synth::make_sample(h->w, s, "", &h->samples[(size_t)s]);
		});
	for (auto &t : pool) t.join();
	return h;
}

tk_handle *tk_load(const char *dir)
{
	tk_handle *h = new tk_handle;
	std::string d = dir;
	std::vector<std::string> bams;
	FILE *fp = fopen((d + "/bamlist").c_str(), "r");
	if (!fp) { h->err = "no bamlist in " + d; return h; }
	char line[1024];
	while (fgets(line, sizeof(line), fp)) {
		size_t n = strlen(line);
		while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
		if (n) bams.push_back(line);
	}
	fclose(fp);
	h->samples.resize(bams.size());
	for (size_t s = 0; s < bams.size(); ++s)
		if (!psc::decode_bam(bams[s], 4, h->samples[s], h->err)) return h;
	if (h->samples.empty()) { h->err = "empty bamlist"; return h; }
	psc::load_subexon_graph(d + "/subexon_combined.out", h->samples[0].chr_names, 0.05, h->g, h->err);
	return h;
}

// writes the graph as a combined-subexon text file, reads it back and compares every array; also the binary cache.
// 0: both round trips reproduce the graph
int tk_graph_roundtrip(tk_handle *h)
{
	char tmpl[] = "/tmp/psc_rt_XXXXXX";
	int fd = mkstemp(tmpl);
	if (fd < 0) return 90;
	close(fd);
	std::string path = tmpl, err;
	int rc = 0;
	auto same = [&](const psc::Graph &a, const psc::Graph &b) {
		return a.n_comp == b.n_comp && a.comp_off == b.comp_off && a.comp_chr == b.comp_chr && a.comp_start == b.comp_start &&
		       a.comp_end == b.comp_end && a.comp_base_gene == b.comp_base_gene && a.comp_used_gene == b.comp_used_gene &&
		       a.se_start == b.se_start && a.se_end == b.se_end && a.se_left_type == b.se_left_type && a.se_right_type == b.se_right_type &&
		       a.se_left_strand == b.se_left_strand && a.se_right_strand == b.se_right_strand && a.se_left_cls == b.se_left_cls &&
		       a.se_right_cls == b.se_right_cls && a.se_gene_id == b.se_gene_id && a.next_off == b.next_off && a.next_idx == b.next_idx &&
		       a.prev_off == b.prev_off && a.prev_idx == b.prev_idx;
	};
	psc::Graph g2, g3;
	if (!psc::write_subexon_file(h->g, path)) rc = 91;
	else if (!psc::load_subexon_graph(path, h->g.chr_names, 0.05, g2, err)) rc = 92;
	else if (!same(h->g, g2)) rc = 93;
	else if (!psc::save_graph_cache(path + ".pscgraph", path, 0.05, g2)) rc = 94;
	else if (!psc::load_graph_cache(path + ".pscgraph", path, h->g.chr_names, 0.05, g3)) rc = 95;
	else if (!same(h->g, g3)) rc = 96;
	else if (psc::load_graph_cache(path + ".pscgraph", path, h->g.chr_names, 0.07, g3)) rc = 97; // other threshold: must be refused
	remove(path.c_str());
	remove((path + ".pscgraph").c_str());
	return rc;
}

int tk_write_subexon_file(tk_handle *h, const char *path) { return psc::write_subexon_file(h->g, path) ? 0 : 1; }

const char *tk_error(tk_handle *h) { return h->err.c_str(); }
void tk_free(tk_handle *h) { delete h; }
int tk_n_samples(tk_handle *h) { return (int)h->samples.size(); }
int tk_n_comp(tk_handle *h) { return h->g.n_comp; }
long long tk_n_reads(tk_handle *h, int s) { return h->samples[(size_t)s].n_reads(); }

void tk_get_graph(tk_handle *h, psc_graph *pg)
{
	const psc::Graph &g = h->g;
	pg->n_comp = g.n_comp; pg->n_se = g.n_se();
	pg->comp_off = g.comp_off.data(); pg->comp_chr = g.comp_chr.data();
	pg->comp_base_gene = g.comp_base_gene.data(); pg->comp_used_gene = g.comp_used_gene.data();
	pg->se_start = g.se_start.data(); pg->se_end = g.se_end.data();
	pg->se_left_type = g.se_left_type.data(); pg->se_right_type = g.se_right_type.data();
	pg->se_left_strand = g.se_left_strand.data(); pg->se_right_strand = g.se_right_strand.data();
	pg->se_left_cls = g.se_left_cls.data(); pg->se_right_cls = g.se_right_cls.data();
	pg->se_gene_id = g.se_gene_id.data();
	pg->next_off = g.next_off.data(); pg->next_idx = g.next_idx.data();
	pg->prev_off = g.prev_off.data(); pg->prev_idx = g.prev_idx.data();
}

void tk_get_reads(tk_handle *h, int s, psc_reads *r)
{
	const psc::SampleReads &sr = h->samples[(size_t)s];
	r->n_reads = sr.n_reads(); r->n_segs = (int64_t)sr.seg_a.size(); r->n_chr = (int)sr.chr_names.size();
	r->chr_off = (const int64_t *)sr.chr_off.data(); r->pos = sr.pos.data(); r->mate_pos = sr.mate_pos.data();
	r->flags = sr.flags.data(); r->nh = sr.nh.data(); r->name_key = sr.name_key.data();
	r->seg_off = sr.seg_off.data(); r->seg_a = sr.seg_a.data(); r->seg_b = sr.seg_b.data();
	r->read_len = sr.read_len; r->frag_stdev = sr.frag_stdev;
}

void tk_get_params(tk_handle *h, psc_params *p)
{
	memset(p, 0, sizeof(*p));
	p->fpkm_fraction = 0.05; p->classifier_threshold = 0.05; p->min_read_depth = 2.5;
	p->max_dp_constraint_size = 7; p->primary_as_unique = 0; p->n_samples = (int)h->samples.size();
	if (!h->samples.empty()) {
		p->read_len = h->samples[0].read_len; p->frag_len = h->samples[0].frag_len;
		p->frag_stdev = h->samples[0].frag_stdev; p->mate_paired = h->samples[0].mate_paired;
	}
}

// writes subexon_combined.out, s<i>.bam and bamlist (synthetic workloads only)
int tk_write_files(tk_handle *h, const char *dir)
{
	if (!h->synthetic) return 1;
	std::string d = dir;
	mkdir(d.c_str(), 0755);
	std::string txt = synth::subexon_text(h->w);
	FILE *fp = fopen((d + "/subexon_combined.out").c_str(), "w");
	if (!fp) return 2;
	fwrite(txt.data(), 1, txt.size(), fp);
	fclose(fp);
	int S = h->w.cfg.n_samples;
	std::vector<std::thread> pool;
	for (int s = 0; s < S; ++s)
		pool.emplace_back([&, s]() { synth::make_sample(h->w, s, d + "/s" + std::to_string(s) + ".bam", nullptr); });
	for (auto &t : pool) t.join();
	fp = fopen((d + "/bamlist").c_str(), "w");
	for (int s = 0; s < S; ++s) fprintf(fp, "%s/s%d.bam\n", d.c_str(), s);
	fclose(fp);
	return 0;
}

int tk_check(tk_handle *h, const char *engine_lib, const CheckOpts *opt, CheckReport *rep)
{
	memset(rep, 0, sizeof(*rep));
	void *lib = dlopen(engine_lib, RTLD_NOW | RTLD_LOCAL);
	if (!lib) { h->err = std::string("dlopen failed: ") + dlerror(); return 100; }
	EngineApi api;
#define SYM(field, name) *(void **)(&api.field) = dlsym(lib, name); if (!api.field) { h->err = std::string("missing symbol ") + name; return 101; }
	SYM(create, "psc_create") SYM(destroy, "psc_destroy") SYM(last_error, "psc_last_error") SYM(set_params, "psc_set_params")
	SYM(set_graph, "psc_set_graph") SYM(set_reads, "psc_set_reads") SYM(build_constraints, "psc_build_constraints")
	SYM(constraint_counts, "psc_constraint_counts") SYM(get_constraint_table, "psc_get_constraint_table") SYM(solve, "psc_solve")
	SYM(transcript_counts, "psc_transcript_counts") SYM(get_transcripts, "psc_get_transcripts") SYM(get_comp_stats, "psc_get_comp_stats")
	SYM(get_timing, "psc_get_timing") SYM(set_coop_width, "psc_set_coop_width")
	CheckOpts o = *opt;
	o.verbose = opt->verbose ? opt->verbose : -1;
	return run_check(h->g, h->samples, api, o, *rep);
}

// Q8: #primary records the reference would have read when the last component ended
double tk_total_read_cnt(tk_handle *h, int s)
{
	const psc::SampleReads &sr = h->samples[(size_t)s];
	const psc::Graph &g = h->g;
	if (g.n_comp == 0) return 0;
	int lastChr = g.comp_chr[(size_t)g.n_comp - 1];
	int maxEnd = 0;
	for (int i = g.comp_off[(size_t)g.n_comp - 1]; i < g.comp_off[(size_t)g.n_comp]; ++i) maxEnd = std::max(maxEnd, g.se_end[(size_t)i]);
	int64_t n = sr.n_reads();
	int64_t a = lastChr < (int)sr.chr_off.size() - 1 ? sr.chr_off[(size_t)lastChr] : n;
	int64_t b = lastChr < (int)sr.chr_off.size() - 1 ? sr.chr_off[(size_t)lastChr + 1] : n;
	int64_t x = a;
	while (x < b && sr.pos[(size_t)x] <= maxEnd) ++x;
	return x < n ? (double)sr.prim_cum[(size_t)x] : (double)sr.total_primary;
}

} // extern "C"
