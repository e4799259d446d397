// CLI of the synthetic generator: writes <out>/subexon_combined.out, <out>/s<i>.bam, <out>/bamlist
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "synth.h"

int main(int argc, char **argv)
{
	synth::Config c;
	std::string out = ".";
	bool writeBam = true;
	for (int a = 1; a < argc; ++a) {
		std::string o = argv[a];
		auto nextd = [&]() { return atof(argv[++a]); };
		auto nexti = [&]() { return atoi(argv[++a]); };
		if (o == "-o") out = argv[++a];
		else if (o == "--seed") c.seed = strtoull(argv[++a], NULL, 10);
		else if (o == "--genes") c.n_genes = nexti();
		else if (o == "--samples") c.n_samples = nexti();
		else if (o == "--chr") c.n_chr = nexti();
		else if (o == "--pairs") c.mean_pairs = nextd();
		else if (o == "--exon-mu") c.exon_mu = nextd();
		else if (o == "--exon-sigma") c.exon_sigma = nextd();
		else if (o == "--exon-min") c.exon_min = nexti();
		else if (o == "--exon-max") c.exon_max = nexti();
		else if (o == "--iso-lambda") c.iso_lambda = nextd();
		else if (o == "--p-skip") c.p_skip = nextd();
		else if (o == "--p-alt") c.p_alt = nextd();
		else if (o == "--p-ir") c.p_ir = nextd();
		else if (o == "--p-noise") c.p_noise = nextd();
		else if (o == "--p-overlap") c.p_overlap_gene = nextd();
		else if (o == "--p-paralog") c.p_paralog = nextd();
		else if (o == "--name-style") c.name_style = nexti();
		else if (o == "--stranded") c.stranded = nexti();
		else if (o == "--dense") c.dense = nexti();
		else if (o == "--graph-noise") c.graph_noise = nexti();
		else if (o == "--single-end") c.single_end = nexti();
		else if (o == "--p-nh3") c.p_nh3 = nextd();
		else if (o == "--frag-sd") c.frag_sd = nextd();
		else if (o == "--no-bam") writeBam = false;
		else { fprintf(stderr, "synth: unknown option %s\n", argv[a]); return 1; }
	}
	mkdir(out.c_str(), 0755);
	synth::Workload w;
	synth::make_genes(c, w);
	{
		std::string txt = synth::subexon_text(w);
		FILE *fp = fopen((out + "/subexon_combined.out").c_str(), "w");
		fwrite(txt.data(), 1, txt.size(), fp);
		fclose(fp);
	}
	if (writeBam) {
		std::vector<std::thread> pool;
		for (int s = 0; s < c.n_samples; ++s)
			pool.emplace_back([&, s]() { synth::make_sample(w, s, out + "/s" + std::to_string(s) + ".bam", nullptr); });
		for (auto &t : pool) t.join();
		FILE *fp = fopen((out + "/bamlist").c_str(), "w");
		for (int s = 0; s < c.n_samples; ++s)
			fprintf(fp, "%s/s%d.bam\n", out.c_str(), s);
		fclose(fp);
	}
	fprintf(stderr, "synth: %d genes, %d samples -> %s\n", (int)w.genes.size(), c.n_samples, out.c_str());
	return 0;
}
