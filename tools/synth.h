// Seeded synthetic stage-3 workload generator (TEST/BENCH INFRASTRUCTURE).
//
// Gene models -> (a) a combined subexon file in the reference's format
// (/root/reference/CombineSubexons.cpp:1617-1684, parsed at SubexonGraph.hpp:275-336) and
// (b) per-sample coordinate-sorted paired-end alignments, either written as BAM files (for the
// reference `classes`) or pushed straight through psc::ReadAccumulator (the same filter the BAM
// decoder applies), so both arms see identical inputs.  Shapes follow SURVEY.md §8(d).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../psiclass_b200/csrc/host/bam_io.h"

namespace synth {

struct Config {
	uint64_t seed = 1;
	int n_chr = 3;
	int n_genes = 60;
	int n_samples = 3;
	double mean_pairs = 100;  // read pairs per gene per sample (before expression factors)
	double exon_mu = 1.6, exon_sigma = 0.9;
	int exon_min = 1, exon_max = 120;
	double iso_lambda = 1.5;  // isoforms per gene = 1 + Poisson
	double p_skip = 0.3, p_alt = 0.1, p_ir = 0.03, p_alt_end = 0.1;
	int read_len = 100;
	double frag_mean = 250, frag_sd = 30;
	double p_nh2 = 0.04, p_nh4 = 0.02;
	double p_nh3 = 0;          // NH=3 reads: 10/3 is inexact, so the order of the weight sum shows (Q9)
	double p_noise = 0.04;     // clipped / intronic / indel / secondary / odd records
	double p_minus = 0.5;
	double p_overlap_gene = 0.05; // next gene placed antisense, overlapping the previous one
	int name_style = 0;        // 0: mates share the id; 1: "/1" "/2" suffixes; 2: mixed incl. ".1" traps
	int stranded = 0;          // 0 un, 1 rf, 2 fr (affects only how flags are drawn)
	double p_paralog = 0;      // fraction of genes duplicated as paralog pairs (NH=2 primary+secondary)
	int dense = 0;             // 1: many exons with several alternative boundaries each (config 4)
	int graph_noise = 1;       // randomise classifiers / add overhang + IR subexons in the direct graph
	int single_end = 0;        // emit unpaired reads only (fragStdev == 0 path)
};

struct Exon { int32_t s, e; }; // 0-based inclusive
struct Isoform { std::vector<Exon> exons; double weight; };
struct Gene {
	int chr, strand;
	int32_t start, end;
	std::vector<Isoform> iso;
	double level;
};

struct Workload {
	Config cfg;
	std::vector<std::string> chr_names;
	std::vector<int32_t> chr_len;
	std::vector<Gene> genes;
};

void make_genes(const Config &cfg, Workload &w);

// the combined subexon file, derived directly from the gene models
std::string subexon_text(const Workload &w);

// alignments of one sample; exactly one of bam_path / reads may be given (or both)
void make_sample(const Workload &w, int sample, const std::string &bam_path, psc::SampleReads *reads);

} // namespace synth
