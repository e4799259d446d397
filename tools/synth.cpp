// See synth.h.  Deterministic for a given Config on every platform (own RNG, no <random>).
#include "synth.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>

namespace synth {
namespace {

struct Rng {
	uint64_t s;
	explicit Rng(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ull + 0x1234567ull) { next(); next(); }
	uint64_t next()
	{
		uint64_t z = (s += 0x9e3779b97f4a7c15ull);
		z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
		z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
		return z ^ (z >> 31);
	}
	double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
	int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); } // inclusive
	double normal()
	{
		double u1 = uni(), u2 = uni();
		if (u1 < 1e-300) u1 = 1e-300;
		return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
	}
	int poisson(double lam)
	{
		if (lam > 50) {
			int v = (int)std::floor(lam + std::sqrt(lam) * normal() + 0.5);
			return v < 0 ? 0 : v;
		}
		double L = std::exp(-lam), p = 1;
		int k = 0;
		do { ++k; p *= uni(); } while (p > L);
		return k - 1;
	}
};

uint64_t mix(uint64_t a, uint64_t b, uint64_t c)
{
	Rng r(a ^ (b * 0xd6e8feb86659fd93ull) ^ (c * 0xca5a826395121157ull));
	return r.next();
}

} // namespace

void make_genes(const Config &cfg, Workload &w)
{
	w.cfg = cfg;
	w.chr_names.clear();
	w.chr_len.clear();
	w.genes.clear();
	w.chr_names.push_back("chrDummy"); // first @SQ: works around GetTrustedSplice.cpp:213 in stage 0
	w.chr_len.push_back(1000);
	for (int c = 0; c < cfg.n_chr; ++c) {
		w.chr_names.push_back("chr" + std::to_string(c + 1));
		w.chr_len.push_back(0);
	}
	Rng rng(cfg.seed);
	std::vector<int32_t> cursor(cfg.n_chr, 5000);
	int prev_chr = -1;
	for (int gidx = 0; gidx < cfg.n_genes; ++gidx) {
		int chr = (int)((int64_t)gidx * cfg.n_chr / cfg.n_genes);
		Gene g;
		g.chr = chr + 1;
		g.strand = rng.uni() < cfg.p_minus ? -1 : 1;
		int n;
		if (cfg.dense)
			n = rng.range(cfg.exon_min > 1 ? cfg.exon_min : 80, cfg.exon_max != 120 ? cfg.exon_max : 200);
		else {
			n = 1 + (int)std::floor(std::exp(cfg.exon_mu + cfg.exon_sigma * rng.normal()));
			n = std::max(cfg.exon_min, std::min(cfg.exon_max, n));
		}
		int32_t pos = cursor[chr];
		bool overlap = prev_chr == chr && !w.genes.empty() && rng.uni() < cfg.p_overlap_gene;
		if (overlap) {
			const Gene &pg = w.genes.back();
			g.strand = -pg.strand;
			pos = std::max<int32_t>(pg.start + 50, pg.end - rng.range(80, 400));
		} else {
			pos += rng.range(1500, 6000);
		}
		// master exons with alternative boundaries
		struct Master { std::vector<int32_t> starts, ends; };
		std::vector<Master> M(n);
		for (int i = 0; i < n; ++i) {
			int len = (i == 0 || i == n - 1) ? rng.range(150, 600) : rng.range(60, 250);
			if (cfg.dense) len = rng.range(40, 160);
			int32_t s = pos, e = pos + len - 1;
			M[i].starts.push_back(s);
			M[i].ends.push_back(e);
			int nalt = cfg.dense ? rng.range(1, 3) : (rng.uni() < cfg.p_alt ? 1 : 0);
			for (int a = 0; a < nalt; ++a) {
				int d = rng.range(8, 30);
				if (rng.uni() < 0.5) {
					if (i > 0 && s + d + 10 < e) M[i].starts.push_back(s + d * (a + 1) < e - 10 ? s + d * (a + 1) : s + d);
				} else {
					if (i < n - 1 && e - d - 10 > s) M[i].ends.push_back(e - d * (a + 1) > s + 10 ? e - d * (a + 1) : e - d);
				}
			}
			pos = e + 1 + rng.range(80, cfg.dense ? 400 : 1500);
		}
		int niso = cfg.dense ? rng.range(20, 60) : 1 + rng.poisson(cfg.iso_lambda);
		std::set<std::vector<int32_t>> seen;
		for (int t = 0; t < niso * 3 && (int)g.iso.size() < niso; ++t) {
			Isoform iso;
			int first = 0, last = n - 1;
			bool canon = g.iso.empty();
			if (!canon && n >= 3) {
				if (rng.uni() < cfg.p_alt_end) first = rng.range(0, std::min(n - 2, 2));
				if (rng.uni() < cfg.p_alt_end) last = rng.range(std::max(first + 1, n - 3), n - 1);
			}
			std::vector<int> keep;
			for (int i = first; i <= last; ++i) {
				if (!canon && i > first && i < last && rng.uni() < cfg.p_skip)
					continue;
				keep.push_back(i);
			}
			for (size_t k = 0; k < keep.size(); ++k) {
				int i = keep[k];
				int32_t s = M[i].starts[0], e = M[i].ends[0];
				if (!canon) {
					if (k > 0) s = M[i].starts[rng.range(0, (int)M[i].starts.size() - 1)];
					if (k + 1 < keep.size()) e = M[i].ends[rng.range(0, (int)M[i].ends.size() - 1)];
				}
				if (s > e) { s = M[i].starts[0]; e = M[i].ends[0]; }
				// intron retention: glue to the previous exon when it is the adjacent master exon
				if (!canon && k > 0 && keep[k - 1] == i - 1 && rng.uni() < cfg.p_ir) {
					iso.exons.back().e = e;
					continue;
				}
				iso.exons.push_back(Exon{s, e});
			}
			std::vector<int32_t> key;
			for (auto &x : iso.exons) { key.push_back(x.s); key.push_back(x.e); }
			if (!seen.insert(key).second)
				continue;
			iso.weight = std::exp(rng.normal());
			g.iso.push_back(iso);
		}
		g.start = M[0].starts[0];
		g.end = M[n - 1].ends[0];
		g.level = std::exp(rng.normal());
		if (!overlap || g.end > cursor[chr])
			cursor[chr] = std::max(cursor[chr], g.end);
		prev_chr = chr;
		w.genes.push_back(g);
		if (cfg.p_paralog > 0 && rng.uni() < cfg.p_paralog) {
			Gene p = g;
			int32_t shift = (g.end - g.start) + rng.range(3000, 8000);
			for (auto &iso : p.iso)
				for (auto &x : iso.exons) { x.s += shift; x.e += shift; }
			p.start += shift;
			p.end += shift;
			p.level = -g.level; // negative: marks "paralog of the previous gene"
			cursor[chr] = p.end;
			w.genes.push_back(p);
		}
	}
	for (int c = 0; c < cfg.n_chr; ++c)
		w.chr_len[c + 1] = cursor[c] + 10000;
}

// -------------------------------------------------------------------------------------------
// direct subexon graph
// -------------------------------------------------------------------------------------------
std::string subexon_text(const Workload &w)
{
	std::string out;
	Rng rng(mix(w.cfg.seed, 77, 1));
	char buf[256];
	// genes are processed in clusters of overlapping genes per chromosome
	size_t gi = 0;
	while (gi < w.genes.size()) {
		size_t gj = gi + 1;
		int32_t cend = w.genes[gi].end;
		while (gj < w.genes.size() && w.genes[gj].chr == w.genes[gi].chr && w.genes[gj].start <= cend + 1) {
			cend = std::max(cend, w.genes[gj].end);
			++gj;
		}
		// boundary marks: coordinate -> (type, strand).  Hard marks win over soft.
		std::map<int32_t, int> cutL, cutR;         // acceptor at s (cut before s) / donor at e (cut after e)
		std::map<int32_t, int> strandL, strandR;
		std::vector<Exon> all;
		std::set<std::pair<int32_t, int32_t>> introns;
		for (size_t k = gi; k < gj; ++k)
			for (auto &iso : w.genes[k].iso)
				for (size_t x = 0; x < iso.exons.size(); ++x) {
					all.push_back(iso.exons[x]);
					if (x > 0) {
						cutL[iso.exons[x].s] = 1;
						strandL[iso.exons[x].s] = w.genes[k].strand;
						introns.insert({iso.exons[x - 1].e, iso.exons[x].s});
					}
					if (x + 1 < iso.exons.size()) {
						cutR[iso.exons[x].e] = 2;
						strandR[iso.exons[x].e] = w.genes[k].strand;
					}
				}
		std::sort(all.begin(), all.end(), [](const Exon &a, const Exon &b) { return a.s < b.s; });
		// union regions
		std::vector<Exon> regions;
		for (auto &e : all) {
			if (!regions.empty() && e.s <= regions.back().e + 1)
				regions.back().e = std::max(regions.back().e, e.e);
			else
				regions.push_back(e);
		}
		struct Piece { int32_t s, e; int lt, rt, ls, rs; double lc, rc; };
		std::vector<Piece> pieces;
		for (auto &r : regions) {
			std::set<int32_t> cuts; // piece starts
			cuts.insert(r.s);
			for (auto it = cutL.lower_bound(r.s + 1); it != cutL.end() && it->first <= r.e; ++it)
				cuts.insert(it->first);
			for (auto it = cutR.lower_bound(r.s); it != cutR.end() && it->first < r.e; ++it)
				cuts.insert(it->first + 1);
			std::vector<int32_t> cs(cuts.begin(), cuts.end());
			for (size_t k = 0; k < cs.size(); ++k) {
				Piece p;
				p.s = cs[k];
				p.e = k + 1 < cs.size() ? cs[k + 1] - 1 : r.e;
				p.lt = cutL.count(p.s) ? 1 : (cutR.count(p.s - 1) ? 2 : 0);
				p.rt = cutR.count(p.e) ? 2 : (cutL.count(p.e + 1) ? 1 : 0);
				// real stage-2 graphs (CombineSubexons) carry a strand and a ~1 classifier only on true
				// splice boundaries (lt==1 / rt==2); internal cuts (lt==2 / rt==1) and soft ends get '.' and -1
				p.ls = p.lt == 1 ? strandL[p.s] : 0;
				p.rs = p.rt == 2 ? strandR[p.e] : 0;
				p.lc = p.lt == 1 ? 1 : -1;
				p.rc = p.rt == 2 ? 1 : -1;
				bool softish = (p.lt == 0 && p.rt == 0) || (p.lt == 0 && p.rt == 1) ||
				               (p.lt == 2 && p.rt == 0) || (p.lt == 2 && p.rt == 1);
				if (softish) {
					// these pieces pass the classifier filter of SubexonGraph.hpp:79-105 only when < threshold
					double v = w.cfg.graph_noise ? (rng.uni() < 0.85 ? rng.uni() * 0.04 : rng.uni()) : 0.0;
					p.lc = p.rc = v;
					p.ls = p.rs = 0;
				} else if (w.cfg.graph_noise) {
					// occasionally a weak hard boundary, which makes it a candidate start/end
					// (TranscriptDecider.cpp:2890-2914)
					if (p.lt == 1 && rng.uni() < 0.06) p.lc = rng.uni() * 0.08;
					else if (p.lt == 1 && rng.uni() < 0.1) p.lc = 0.9 + 0.1 * rng.uni();
					if (p.rt == 2 && rng.uni() < 0.06) p.rc = rng.uni() * 0.08;
					else if (p.rt == 2 && rng.uni() < 0.1) p.rc = 0.9 + 0.1 * rng.uni();
				}
				pieces.push_back(p);
			}
		}
		// edges: adjacency + introns (coordinates, 1-based)
		std::map<int32_t, size_t> byStart, byEnd;
		for (size_t k = 0; k < pieces.size(); ++k) {
			byStart[pieces[k].s] = k;
			byEnd[pieces[k].e] = k;
		}
		std::vector<std::vector<int32_t>> prevC(pieces.size()), nextC(pieces.size());
		for (size_t k = 0; k + 1 < pieces.size(); ++k)
			if (pieces[k].e + 1 == pieces[k + 1].s) {
				nextC[k].push_back(pieces[k + 1].s + 1);
				prevC[k + 1].push_back(pieces[k].e + 1);
			}
		for (auto &in : introns) {
			auto a = byEnd.find(in.first);
			auto b = byStart.find(in.second);
			if (a == byEnd.end() || b == byStart.end())
				continue;
			nextC[a->second].push_back(in.second + 1);
			prevC[b->second].push_back(in.first + 1);
		}
		const std::string &chr = w.chr_names[w.genes[gi].chr];
		for (size_t k = 0; k < pieces.size(); ++k) {
			Piece &p = pieces[k];
			std::sort(prevC[k].begin(), prevC[k].end());
			std::sort(nextC[k].begin(), nextC[k].end());
			prevC[k].erase(std::unique(prevC[k].begin(), prevC[k].end()), prevC[k].end());
			nextC[k].erase(std::unique(nextC[k].begin(), nextC[k].end()), nextC[k].end());
			const char *sc = "-.+";
			snprintf(buf, sizeof(buf), "%s %d %d %d %d %c %c -1 -1 -1 %.6f %.6f %d ", chr.c_str(), p.s + 1,
			         p.e + 1, p.lt, p.rt, sc[p.ls + 1], sc[p.rs + 1], p.lc, p.rc, (int)prevC[k].size());
			out += buf;
			for (int32_t v : prevC[k]) { snprintf(buf, sizeof(buf), "%d ", v); out += buf; }
			snprintf(buf, sizeof(buf), "%d ", (int)nextC[k].size());
			out += buf;
			for (int32_t v : nextC[k]) { snprintf(buf, sizeof(buf), "%d ", v); out += buf; }
			out += "\n";
		}
		gi = gj;
	}
	return out;
}

// -------------------------------------------------------------------------------------------
// reads
// -------------------------------------------------------------------------------------------
namespace {

struct Rec {
	int32_t pos, mpos, tlen, l_seq;
	int16_t tid, mtid;
	uint16_t flag;
	int8_t nh;   // -1: no NH tag
	char xs;     // 0: none
	uint8_t ncig;
	uint32_t cig[14];
	uint32_t name_id; // index into a per-call name table
	uint32_t order;
};

// blocks of transcript interval [t0, t0+len) on the genome
int map_blocks(const Isoform &iso, int32_t t0, int32_t len, int32_t *bs, int32_t *be, int maxb)
{
	int nb = 0;
	int32_t acc = 0;
	for (auto &x : iso.exons) {
		int32_t xl = x.e - x.s + 1;
		int32_t lo = std::max(t0, acc), hi = std::min(t0 + len, acc + xl);
		if (lo < hi) {
			int32_t gs = x.s + (lo - acc), ge = x.s + (hi - acc) - 1;
			if (nb > 0 && be[nb - 1] + 1 == gs)
				be[nb - 1] = ge;
			else {
				if (nb >= maxb) return -1;
				bs[nb] = gs;
				be[nb] = ge;
				++nb;
			}
		}
		acc += xl;
		if (acc >= t0 + len) break;
	}
	return nb;
}

bool fill_cigar(Rec &r, const int32_t *bs, const int32_t *be, int nb, int headClip, int tailClip, int ins, bool hardHead)
{
	int k = 0;
	if (headClip > 0) r.cig[k++] = ((uint32_t)headClip << 4) | (hardHead ? 5u : 4u);
	for (int b = 0; b < nb; ++b) {
		if (k + 4 >= 14) return false;
		int32_t len = be[b] - bs[b] + 1;
		if (b > 0) r.cig[k++] = ((uint32_t)(bs[b] - be[b - 1] - 1) << 4) | 3u;
		if (b == 0 && ins > 0 && len > 20) {
			r.cig[k++] = (10u << 4) | 0u;
			r.cig[k++] = ((uint32_t)ins << 4) | 1u;
			r.cig[k++] = ((uint32_t)(len - 10) << 4) | 0u;
		} else
			r.cig[k++] = ((uint32_t)len << 4) | 0u;
	}
	if (tailClip > 0) {
		if (k >= 14) return false;
		r.cig[k++] = ((uint32_t)tailClip << 4) | 4u;
	}
	r.ncig = (uint8_t)k;
	return true;
}

} // namespace

void make_sample(const Workload &w, int sample, const std::string &bam_path, psc::SampleReads *reads)
{
	const Config &cfg = w.cfg;
	psc::BamWriter bw;
	bool toBam = !bam_path.empty();
	if (toBam)
		bw.open(bam_path, w.chr_names, w.chr_len);
	psc::ReadAccumulator *acc = nullptr;
	if (reads) {
		*reads = psc::SampleReads();
		reads->path = bam_path;
		reads->chr_names = w.chr_names;
		reads->chr_len = w.chr_len;
		acc = new psc::ReadAccumulator(*reads);
	}
	std::vector<Rec> recs;
	std::vector<std::string> names;
	auto flush = [&]() {
		std::stable_sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
			if (a.pos != b.pos) return a.pos < b.pos;
			return a.order < b.order;
		});
		std::string aux;
		for (auto &r : recs) {
			const std::string &nm = names[r.name_id];
			if (toBam) {
				aux.clear();
				if (r.nh >= 0) { aux += "NHC"; aux += (char)r.nh; }
				if (r.xs) { aux += "XSA"; aux += r.xs; }
				bw.write(r.tid, r.pos, r.flag, 60, nm, r.cig, r.ncig, r.l_seq, r.mtid, r.mpos, r.tlen, aux);
			}
			if (acc)
				acc->add(r.tid, r.pos, r.flag, r.l_seq, r.mtid, r.mpos, nm.data(), nm.size(), r.cig, r.ncig,
				         r.nh >= 0, r.nh);
		}
		recs.clear();
		names.clear();
	};
	double sampleFactor = std::exp(0.3 * Rng(mix(cfg.seed, 1000 + sample, 5)).normal());
	int cur_chr = -1;
	int32_t cluster_end = -1;
	uint32_t order = 0;
	for (size_t gidx = 0; gidx < w.genes.size(); ++gidx) {
		const Gene &g = w.genes[gidx];
		if (g.chr != cur_chr || g.start > cluster_end + 2000) {
			flush();
			cur_chr = g.chr;
			cluster_end = g.end;
		}
		cluster_end = std::max(cluster_end, g.end);
		if (g.level < 0)
			continue; // paralog copy: its reads are emitted with the original
		const Gene *para = (gidx + 1 < w.genes.size() && w.genes[gidx + 1].level < 0) ? &w.genes[gidx + 1] : nullptr;
		if (para) cluster_end = std::max(cluster_end, para->end);
		Rng rng(mix(cfg.seed, (uint64_t)sample + 1, gidx + 1));
		// per-sample isoform usage
		std::vector<double> wt(g.iso.size());
		double wsum = 0;
		for (size_t i = 0; i < g.iso.size(); ++i) {
			double f = rng.uni() < 0.15 ? 0.0 : std::exp(0.7 * rng.normal());
			int32_t tl = 0;
			for (auto &x : g.iso[i].exons) tl += x.e - x.s + 1;
			wt[i] = g.iso[i].weight * f * tl;
			wsum += wt[i];
		}
		if (wsum <= 0) { wt[0] = 1; wsum = 1; }
		double gf = std::exp(0.5 * rng.normal());
		int npairs = rng.poisson(std::min(cfg.mean_pairs * g.level * gf * sampleFactor, cfg.mean_pairs * 40));
		for (int p = 0; p < npairs; ++p) {
			double u = rng.uni() * wsum;
			size_t ii = 0;
			while (ii + 1 < wt.size() && u >= wt[ii]) { u -= wt[ii]; ++ii; }
			const Isoform &iso = g.iso[ii];
			int32_t tl = 0;
			for (auto &x : iso.exons) tl += x.e - x.s + 1;
			int rl = cfg.read_len;
			if (tl < rl) continue;
			int frag = (int)std::floor(cfg.frag_mean + cfg.frag_sd * rng.normal() + 0.5);
			frag = std::max(rl, std::min(frag, (int)tl));
			int noise = rng.uni() < cfg.p_noise ? rng.range(0, 9) : -1;
			if (noise == 8) frag = rl; // both mates at the same position
			int32_t t0 = (int32_t)(rng.uni() * (tl - frag + 1));
			if (t0 > tl - frag) t0 = tl - frag;
			int nh = 1;
			double un = rng.uni();
			if (un < cfg.p_nh4) nh = 4;
			else if (un < cfg.p_nh4 + cfg.p_nh2) nh = 2;
			else if (un < cfg.p_nh4 + cfg.p_nh2 + cfg.p_nh3) nh = 3;
			if (noise == 6) nh = 6;
			if (para) nh = 2;
			bool leftFirst;
			if (cfg.stranded == 1) leftFirst = g.strand < 0;
			else if (cfg.stranded == 2) leftFirst = g.strand > 0;
			else leftFirst = rng.uni() < 0.5;

			Rec a, b;
			memset(&a, 0, sizeof(a));
			memset(&b, 0, sizeof(b));
			int32_t bs[16], be[16];
			int headClip = noise == 0 ? rng.range(3, 25) : (noise == 7 ? 10 : 0);
			int tailClip = noise == 1 ? rng.range(3, 25) : 0;
			int ins = noise == 3 ? 3 : 0;
			// mate A (left)
			int nb;
			if (noise == 2) { // unspliced, runs into the intron
				int32_t gs[1], ge[1];
				int nbx = map_blocks(iso, t0, 1, gs, ge, 1);
				if (nbx != 1) continue;
				bs[0] = gs[0];
				be[0] = gs[0] + rl - 1;
				nb = 1;
			} else
				nb = map_blocks(iso, t0 + headClip, rl - headClip - ins, bs, be, 12);
			if (nb <= 0) continue;
			a.l_seq = (noise == 7) ? rl - headClip : rl;
			if (!fill_cigar(a, bs, be, nb, headClip, 0, ins, noise == 7)) continue;
			a.pos = bs[0];
			int32_t aEnd = be[nb - 1];
			a.xs = nb > 1 ? (g.strand > 0 ? '+' : '-') : 0;
			// mate B (right)
			nb = map_blocks(iso, t0 + frag - rl, rl - tailClip, bs, be, 12);
			if (nb <= 0) continue;
			b.l_seq = rl;
			if (!fill_cigar(b, bs, be, nb, 0, tailClip, 0, false)) continue;
			b.pos = bs[0];
			int32_t bEnd = be[nb - 1];
			b.xs = nb > 1 ? (g.strand > 0 ? '+' : '-') : 0;
			(void)aEnd;

			a.tid = b.tid = (int16_t)g.chr;
			a.nh = b.nh = (int8_t)nh;
			if (noise == 9) a.nh = b.nh = -1; // no NH tag at all
			std::string base = "g" + std::to_string(gidx) + "s" + std::to_string(sample) + "p" + std::to_string(p);
			std::string na = base, nbm = base;
			int style = cfg.name_style;
			if (style == 2) {
				int t = rng.range(0, 5);
				style = t < 2 ? 0 : (t < 4 ? 1 : (t == 4 ? 3 : 4));
			}
			if (style == 1) { na += leftFirst ? "/1" : "/2"; nbm += leftFirst ? "/2" : "/1"; }
			else if (style == 3) { na += leftFirst ? ".1" : ".2"; nbm += leftFirst ? ".2" : ".1"; }
			else if (style == 4) { na += ".1"; nbm += ".1"; } // shared id that LOOKS suffixed: Q2 breaks the pairing
			if (cfg.single_end) {
				a.flag = rng.uni() < 0.5 ? 0 : 0x10;
				a.mtid = -1; a.mpos = -1; a.tlen = 0;
				a.name_id = (uint32_t)names.size(); names.push_back(na);
				a.order = order++;
				recs.push_back(a);
				continue;
			}
			a.flag = 0x1 | 0x2 | 0x20 | (leftFirst ? 0x40 : 0x80);
			b.flag = 0x1 | 0x2 | 0x10 | (leftFirst ? 0x80 : 0x40);
			a.mtid = b.mtid = (int16_t)g.chr;
			a.mpos = b.pos;
			b.mpos = a.pos;
			a.tlen = bEnd - a.pos + 1;
			b.tlen = -a.tlen;
			if (noise == 5) { // mate B unmapped
				a.flag = (uint16_t)((a.flag & ~0x2) | 0x8);
				a.mpos = a.pos;
				b.flag = (uint16_t)(0x1 | 0x4 | (leftFirst ? 0x80 : 0x40));
				b.pos = a.pos;
				b.mpos = a.pos;
				b.ncig = 0;
				b.tlen = a.tlen = 0;
				b.xs = 0;
			}
			a.name_id = (uint32_t)names.size(); names.push_back(na);
			b.name_id = (uint32_t)names.size(); names.push_back(nbm);
			a.order = order++;
			b.order = order++;
			recs.push_back(a);
			recs.push_back(b);
			if (noise == 4 || para) { // secondary copy of mate A (at the paralog when there is one)
				Rec s = a;
				s.flag |= 0x100;
				if (noise == 4 && !para) { recs[recs.size() - 2].nh = 2; s.nh = 2; }
				if (para) {
					int32_t shift = para->start - g.start;
					s.pos += shift;
					s.mpos += shift;
					Rec s2 = b;
					s2.flag |= 0x100;
					s2.pos += shift;
					s2.mpos += shift;
					s2.order = order++;
					recs.push_back(s2);
				}
				s.order = order++;
				recs.push_back(s);
			}
		}
	}
	flush();
	if (toBam)
		bw.close();
	if (acc) {
		acc->finish();
		delete acc;
	}
}

} // namespace synth
