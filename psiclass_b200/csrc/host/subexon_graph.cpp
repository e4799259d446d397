// See subexon_graph.h.  Every step cites the reference lines whose behaviour it reproduces.
#include "subexon_graph.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace psc {
namespace {

struct RawSubexon {
	int32_t chr, start, end; // 1-based while parsing (SubexonGraph.hpp:166-172 shifts later)
	int32_t lt, rt, ls, rs;
	double lcls, rcls;
	std::vector<int32_t> prev, next; // coordinates first, then global indices
};

inline bool same_strand(int a, int b) { return a == 0 || b == 0 || a == b; }

// one line of the combined subexon file (SubexonGraph.hpp:275-336; format written by
// CombineSubexons.cpp:1617-1684)
bool parse_line(char *line, const std::map<std::string, int> &chr_id, RawSubexon &se, std::string &err)
{
	std::vector<char *> tok;
	for (char *p = strtok(line, " \t\r\n"); p; p = strtok(NULL, " \t\r\n"))
		tok.push_back(p);
	if (tok.size() < 14) {
		err = "malformed subexon line";
		return false;
	}
	auto it = chr_id.find(tok[0]);
	if (it == chr_id.end()) {
		err = std::string("Unknown genome name: ") + tok[0];
		return false;
	}
	se.chr = it->second;
	se.start = atoi(tok[1]);
	se.end = atoi(tok[2]);
	se.lt = atoi(tok[3]);
	se.rt = atoi(tok[4]);
	se.ls = tok[5][0] == '+' ? 1 : tok[5][0] == '-' ? -1 : 0;
	se.rs = tok[6][0] == '+' ? 1 : tok[6][0] == '-' ? -1 : 0;
	se.lcls = strtod(tok[10], NULL);
	se.rcls = strtod(tok[11], NULL);
	size_t k = 12;
	int pc = atoi(tok[k++]);
	se.prev.clear();
	for (int i = 0; i < pc && k < tok.size(); ++i)
		se.prev.push_back(atoi(tok[k++]));
	int nc = k < tok.size() ? atoi(tok[k++]) : 0;
	se.next.clear();
	for (int i = 0; i < nc && k < tok.size(); ++i)
		se.next.push_back(atoi(tok[k++]));
	return true;
}

struct GeneIdState {
	const std::vector<RawSubexon> *all;
	int base; // first global index of the component
	int n;
	std::vector<int32_t> *gid;                   // local
	const std::vector<std::vector<int32_t>> *nx; // local adjacency
	const std::vector<std::vector<int32_t>> *pv;
};

// SubexonGraph::SetGeneId (SubexonGraph.cpp:151-200)
void set_gene_id(GeneIdState &st, int tag, int strand, int id)
{
	std::vector<int32_t> &gid = *st.gid;
	const std::vector<RawSubexon> &S = *st.all;
	if (gid[tag] != -1 && gid[tag] != -2) {
		if (gid[tag] != id)
			gid[tag] = -2;
		else
			return;
	} else if (gid[tag] == -2)
		return;
	if (gid[tag] != -2)
		gid[tag] = id;
	const RawSubexon &me = S[st.base + tag];
	if (same_strand(strand, me.rs)) {
		for (int nb : (*st.nx)[tag])
			if (S[st.base + nb].start > me.end + 1)
				set_gene_id(st, nb, strand, id);
	}
	if (same_strand(strand, me.ls)) {
		for (int nb : (*st.pv)[tag])
			if (S[st.base + nb].end < me.start - 1)
				set_gene_id(st, nb, strand, id);
	}
	if (tag < st.n - 1 && S[st.base + tag + 1].start == me.end + 1)
		set_gene_id(st, tag + 1, strand, id);
	if (tag > 0 && me.start - 1 == S[st.base + tag - 1].end)
		set_gene_id(st, tag - 1, strand, id);
}

// SubexonGraph::GetGeneBoundary (SubexonGraph.cpp:3-20), iterative
void gene_boundary(const std::vector<RawSubexon> &S, std::vector<int> &visit, int tag, int &boundary,
                   int stamp, std::vector<int> &stack)
{
	stack.clear();
	stack.push_back(tag);
	while (!stack.empty()) {
		int t = stack.back();
		stack.pop_back();
		if (visit[t] == stamp)
			continue;
		visit[t] = stamp;
		if (S[t].end > boundary)
			boundary = S[t].end;
		for (size_t i = S[t].next.size(); i-- > 0;)
			stack.push_back(S[t].next[i]);
	}
}

} // namespace

bool load_subexon_graph(const std::string &path, const std::vector<std::string> &chr_names,
                        double thr, Graph &g, std::string &err)
{
	FILE *fp = fopen(path.c_str(), "r");
	if (!fp) {
		err = "Cannot find combined subexon file.";
		return false;
	}
	std::map<std::string, int> chr_id;
	for (size_t i = 0; i < chr_names.size(); ++i)
		chr_id[chr_names[i]] = (int)i;

	// ---- read + filter (SubexonGraph.hpp:70-109) ----
	std::vector<RawSubexon> S;
	{
		std::vector<char> buf(1 << 16);
		std::string line;
		while (fgets(buf.data(), (int)buf.size(), fp)) {
			line = buf.data();
			while (!line.empty() && line.back() != '\n' && fgets(buf.data(), (int)buf.size(), fp))
				line += buf.data();
			if (line.empty() || line[0] == '#')
				continue;
			RawSubexon se;
			std::vector<char> tmp(line.begin(), line.end());
			tmp.push_back(0);
			if (!parse_line(tmp.data(), chr_id, se, err)) {
				fclose(fp);
				return false;
			}
			bool softish = (se.lt == 0 && se.rt == 0) || (se.lt == 0 && se.rt == 1) ||
			               (se.lt == 2 && se.rt == 0) || (se.lt == 2 && se.rt == 1);
			if (softish) {
				bool overhang = (se.lt == 0 && se.rt == 1) || (se.lt == 2 && se.rt == 0);
				if (overhang && se.end - se.start + 1 <= 7)
					continue;
				if (se.lcls >= thr || se.lcls < 0)
					continue;
			}
			S.push_back(std::move(se));
		}
		fclose(fp);
	}
	int n = (int)S.size();

	// ---- coordinates -> indices (SubexonGraph.hpp:113-164) ----
	for (int i = 0; i < n; ++i) {
		RawSubexon &se = S[i];
		int pc = (int)se.prev.size();
		int cnt = 0, bound = 0, j, k;
		if (pc > 0)
			bound = se.prev[0];
		for (j = i - 1, k = 0; k < pc && j >= 0 && S[j].end >= bound; --j) {
			if (S[j].end == se.prev[pc - 1 - k]) {
				se.prev[pc - 1 - cnt] = j;
				++k;
				++cnt;
			} else if (S[j].end < se.prev[pc - 1 - k]) {
				++k;
				++j;
			}
		}
		for (j = 0, k = pc - cnt; j < cnt; ++j, ++k)
			se.prev[j] = se.prev[k];
		se.prev.resize(cnt);

		int nc = (int)se.next.size();
		cnt = 0;
		if (nc > 0)
			bound = se.next[nc - 1];
		for (j = i + 1, k = 0; k < nc && j < n && S[j].start <= bound; ++j) {
			if (S[j].start == se.next[k]) {
				se.next[cnt] = j;
				++k;
				++cnt;
			} else if (S[j].start > se.next[k]) {
				++k;
				--j;
			}
		}
		se.next.resize(cnt);
	}
	for (int i = 0; i < n; ++i) {
		--S[i].start;
		--S[i].end;
	}

	// ---- classifier of hard boundaries next to an overhang (SubexonGraph.hpp:176-194) ----
	for (int i = 0; i < n; ++i) {
		int j;
		if (S[i].lt == 1 && S[i].lcls < 1) {
			for (j = i - 1; j >= 0; --j)
				if (S[j].end < S[j + 1].start - 1)
					break;
			if (S[j + 1].lt == 0)
				S[i].lcls = 1;
		}
		if (S[i].rt == 2 && S[i].rcls < 1) {
			for (j = i + 1; j < n; ++j)
				if (S[j].start > S[j - 1].end + 1)
					break;
			if (S[j - 1].rt == 0)
				S[i].rcls = 1;
		}
	}

	// ---- mixed-strand regions: force terminal sites (SubexonGraph.hpp:198-259) ----
	for (int i = 0; i < n;) {
		int j, k;
		int support[2] = {0, 0};
		for (j = i + 1; j < n; ++j)
			if (S[j].start > S[j - 1].end + 1)
				break;
		for (k = i; k < j; ++k) {
			if (S[k].ls != 0) ++support[(S[k].ls + 1) / 2];
			if (S[k].rs != 0) ++support[(S[k].rs + 1) / 2];
		}
		if (support[0] == 0 || support[1] == 0) {
			i = j;
			continue;
		}
		int leftSupport[2] = {0, 0}, rightSupport[2] = {0, 0};
		for (k = i; k < j; ++k) {
			if (S[k].ls != 0)
				for (int p : S[k].prev)
					if (p < i) {
						++leftSupport[(S[k].ls + 1) / 2];
						break;
					}
			if (S[k].rs != 0)
				for (int q : S[k].next)
					if (q >= j) {
						++rightSupport[(S[k].rs + 1) / 2];
						break;
					}
		}
		if (((leftSupport[0] > 0 && rightSupport[0] == 0) || (leftSupport[1] > 0 && rightSupport[1] == 0)) &&
		    S[j - 1].rt != 0)
			S[j - 1].rcls = 0;
		if (((leftSupport[0] == 0 && rightSupport[0] > 0) || (leftSupport[1] == 0 && rightSupport[1] > 0)) &&
		    S[j - 1].lt != 0)
			S[j - 1].lcls = 0;
		i = j;
	}

	// ---- gene intervals (SubexonGraph.cpp:22-112) ----
	g = Graph();
	g.chr_names = chr_names;
	std::vector<int> visit(n, -1), stack;
	struct Interval { int startIdx, endIdx, start, end; };
	std::vector<Interval> intervals;
	{
		int tag = 0, cnt = 0;
		while (tag < n) {
			Interval gi;
			int farthest = -1;
			gene_boundary(S, visit, tag, farthest, cnt, stack);
			int i;
			for (i = tag + 1; i < n; ++i) {
				if (S[i].start > farthest || S[i].chr != S[tag].chr)
					break;
				gene_boundary(S, visit, i, farthest, cnt, stack);
			}
			gi.endIdx = i - 1;
			++cnt;
			gi.startIdx = tag;
			gi.start = S[gi.startIdx].start;
			gi.end = S[gi.endIdx].end;
			tag = gi.endIdx + 1;
			if (S[gi.startIdx].ls != 0 && S[gi.startIdx].ls != S[gi.startIdx].rs) {
				for (i = gi.startIdx; i >= 0; --i)
					if ((S[i].lt == 1 && S[i].lcls < thr) || S[i].lt == 0 ||
					    (i > 0 && S[i - 1].end + 1 < S[i].start))
						break;
				gi.start = S[i < 0 ? 0 : i].start; // i reaches -1 only if S[0] fails every test (then the reference reads S[-1])
			}
			if (S[gi.endIdx].rs != 0 && S[gi.endIdx].ls != S[gi.endIdx].rs) {
				for (i = gi.endIdx; i < n; ++i)
					if ((S[i].rt == 2 && S[i].rcls < thr) || S[i].rt == 0 ||
					    (i < n - 1 && S[i].end + 1 < S[i + 1].start))
						break;
				gi.end = S[i >= n ? n - 1 : i].end;
			}
			intervals.push_back(gi);
		}
	}

	// ---- per-component extraction + sequential gene ids (SubexonGraph.cpp:114-241) ----
	g.n_comp = (int)intervals.size();
	g.comp_off.assign(g.n_comp + 1, 0);
	g.se_start.resize(n);
	g.se_end.resize(n);
	g.se_left_type.resize(n);
	g.se_right_type.resize(n);
	g.se_left_strand.resize(n);
	g.se_right_strand.resize(n);
	g.se_left_cls.resize(n);
	g.se_right_cls.resize(n);
	g.se_gene_id.resize(n);
	g.next_off.assign(n + 1, 0);
	g.prev_off.assign(n + 1, 0);
	int usedGeneId = 0;
	for (int c = 0; c < g.n_comp; ++c) {
		const Interval &gi = intervals[c];
		int base = gi.startIdx, cnt = gi.endIdx - gi.startIdx + 1;
		g.comp_off[c] = base;
		g.comp_chr.push_back(S[base].chr);
		g.comp_start.push_back(gi.start);
		g.comp_end.push_back(gi.end);
		std::vector<std::vector<int32_t>> nx(cnt), pv(cnt);
		for (int i = 0; i < cnt; ++i) {
			for (int p : S[base + i].prev)
				if (p - base >= 0 && p - base < cnt)
					pv[i].push_back(p - base);
			for (int q : S[base + i].next)
				if (q - base >= 0 && q - base < cnt)
					nx[i].push_back(q - base);
		}
		std::vector<int32_t> gid(cnt, -1);
		GeneIdState st{&S, base, cnt, &gid, &nx, &pv};
		int baseGeneId = usedGeneId;
		int lastMinus = -1;
		for (int strand = -1; strand <= 1; strand += 2)
			for (int i = 0; i < cnt; ++i) {
				const RawSubexon &se = S[base + i];
				if ((gid[i] == -1 && ((strand == 1 && se.rs == 0) || se.rs == strand)) ||
				    (strand == 1 && baseGeneId <= gid[i] && gid[i] <= lastMinus && se.rs == strand)) {
					set_gene_id(st, i, strand, usedGeneId);
					if (strand == -1)
						lastMinus = usedGeneId;
					++usedGeneId;
				}
			}
		for (int i = 0; i < cnt; ++i)
			if (S[base + i].lt == 0 && S[base + i].rt == 0)
				gid[i] = usedGeneId++;
		g.comp_base_gene.push_back(baseGeneId);
		g.comp_used_gene.push_back(usedGeneId);
		for (int i = 0; i < cnt; ++i) {
			const RawSubexon &se = S[base + i];
			int gi2 = base + i;
			g.se_start[gi2] = se.start;
			g.se_end[gi2] = se.end;
			g.se_left_type[gi2] = (int8_t)se.lt;
			g.se_right_type[gi2] = (int8_t)se.rt;
			g.se_left_strand[gi2] = (int8_t)se.ls;
			g.se_right_strand[gi2] = (int8_t)se.rs;
			g.se_left_cls[gi2] = se.lcls;
			g.se_right_cls[gi2] = se.rcls;
			g.se_gene_id[gi2] = gid[i];
			for (int q : nx[i]) g.next_idx.push_back(q);
			for (int p : pv[i]) g.prev_idx.push_back(p);
			g.next_off[gi2 + 1] = (int32_t)g.next_idx.size();
			g.prev_off[gi2 + 1] = (int32_t)g.prev_idx.size();
		}
	}
	g.comp_off[g.n_comp] = n;
	return true;
}

// -------------------------------------------------------------------------------------------
// text writer
// -------------------------------------------------------------------------------------------
bool write_subexon_file(const Graph &g, const std::string &path)
{
	FILE *fp = fopen(path.c_str(), "w");
	if (!fp) return false;
	static const char sc[] = "-.+";
	for (int c = 0; c < g.n_comp; ++c) {
		const int b = g.comp_off[(size_t)c], e = g.comp_off[(size_t)c + 1];
		const char *chr = g.chr_names[(size_t)g.comp_chr[(size_t)c]].c_str();
		for (int i = b; i < e; ++i) {
			fprintf(fp, "%s %d %d %d %d %c %c -1 -1 -1 %.17g %.17g ", chr, g.se_start[(size_t)i] + 1, g.se_end[(size_t)i] + 1,
			        (int)g.se_left_type[(size_t)i], (int)g.se_right_type[(size_t)i], sc[g.se_left_strand[(size_t)i] + 1],
			        sc[g.se_right_strand[(size_t)i] + 1], g.se_left_cls[(size_t)i], g.se_right_cls[(size_t)i]);
			fprintf(fp, "%d ", g.prev_off[(size_t)i + 1] - g.prev_off[(size_t)i]);
			for (int k = g.prev_off[(size_t)i]; k < g.prev_off[(size_t)i + 1]; ++k) fprintf(fp, "%d ", g.se_end[(size_t)(b + g.prev_idx[(size_t)k])] + 1);
			fprintf(fp, "%d ", g.next_off[(size_t)i + 1] - g.next_off[(size_t)i]);
			for (int k = g.next_off[(size_t)i]; k < g.next_off[(size_t)i + 1]; ++k) fprintf(fp, "%d ", g.se_start[(size_t)(b + g.next_idx[(size_t)k])] + 1);
			fprintf(fp, "\n");
		}
	}
	return fclose(fp) == 0;
}

// -------------------------------------------------------------------------------------------
// binary cache
// -------------------------------------------------------------------------------------------
namespace {
const uint64_t CACHE_MAGIC = 0x3148504152474350ull; // "PCGRAPH1"

bool source_stamp(const std::string &path, uint64_t &size, uint64_t &mtime)
{
	struct stat st;
	if (stat(path.c_str(), &st) != 0) return false;
	size = (uint64_t)st.st_size;
	mtime = (uint64_t)st.st_mtime;
	return true;
}
template <class T> void put_vec(std::string &out, const std::vector<T> &v)
{
	uint64_t n = v.size();
	out.append((const char *)&n, 8);
	if (n) out.append((const char *)v.data(), n * sizeof(T));
	out.append((8 - out.size() % 8) % 8, '\0');
}
template <class T> bool get_vec(const char *&p, const char *end, std::vector<T> &v)
{
	if (end - p < 8) return false;
	uint64_t n;
	memcpy(&n, p, 8);
	p += 8;
	if ((uint64_t)(end - p) < n * sizeof(T)) return false;
	v.resize((size_t)n);
	if (n) memcpy(v.data(), p, (size_t)n * sizeof(T));
	p += n * sizeof(T);
	p += (8 - (uintptr_t)(n * sizeof(T)) % 8) % 8;
	return p <= end;
}
} // namespace

bool save_graph_cache(const std::string &cachePath, const std::string &sourcePath, double thr, const Graph &g)
{
	uint64_t size = 0, mtime = 0;
	if (!source_stamp(sourcePath, size, mtime)) return false;
	std::string out;
	out.append((const char *)&CACHE_MAGIC, 8);
	out.append((const char *)&size, 8);
	out.append((const char *)&mtime, 8);
	out.append((const char *)&thr, 8);
	uint64_t nchr = g.chr_names.size();
	out.append((const char *)&nchr, 8);
	for (auto &c : g.chr_names) {
		uint64_t l = c.size();
		out.append((const char *)&l, 8);
		out.append(c);
		out.append((8 - out.size() % 8) % 8, '\0');
	}
	int64_t ncomp = g.n_comp;
	out.append((const char *)&ncomp, 8);
	put_vec(out, g.comp_off); put_vec(out, g.comp_chr); put_vec(out, g.comp_start); put_vec(out, g.comp_end);
	put_vec(out, g.comp_base_gene); put_vec(out, g.comp_used_gene);
	put_vec(out, g.se_start); put_vec(out, g.se_end); put_vec(out, g.se_left_type); put_vec(out, g.se_right_type);
	put_vec(out, g.se_left_strand); put_vec(out, g.se_right_strand); put_vec(out, g.se_left_cls); put_vec(out, g.se_right_cls);
	put_vec(out, g.se_gene_id); put_vec(out, g.next_off); put_vec(out, g.next_idx); put_vec(out, g.prev_off); put_vec(out, g.prev_idx);
	std::string tmp = cachePath + ".tmp";
	FILE *fp = fopen(tmp.c_str(), "wb");
	if (!fp) return false;
	bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
	ok = fclose(fp) == 0 && ok;
	if (ok) ok = rename(tmp.c_str(), cachePath.c_str()) == 0; // readers never see a half-written cache
	if (!ok) remove(tmp.c_str());
	return ok;
}

bool load_graph_cache(const std::string &cachePath, const std::string &sourcePath, const std::vector<std::string> &chrNames, double thr,
                      Graph &g)
{
	uint64_t size = 0, mtime = 0;
	if (!source_stamp(sourcePath, size, mtime)) return false;
	int fd = open(cachePath.c_str(), O_RDONLY);
	if (fd < 0) return false;
	struct stat st;
	if (fstat(fd, &st) != 0 || st.st_size < 48) {
		close(fd);
		return false;
	}
	void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (map == MAP_FAILED) return false;
	const char *p = (const char *)map, *end = p + st.st_size;
	Graph t;
	bool ok = false;
	do {
		uint64_t hdr[5];
		memcpy(hdr, p, 40);
		p += 40;
		double cthr;
		memcpy(&cthr, &hdr[3], 8);
		if (hdr[0] != CACHE_MAGIC || hdr[1] != size || hdr[2] != mtime || cthr != thr || hdr[4] != chrNames.size()) break;
		bool same = true;
		for (size_t k = 0; k < chrNames.size() && same; ++k) {
			if (end - p < 8) { same = false; break; }
			uint64_t l;
			memcpy(&l, p, 8);
			p += 8;
			if ((uint64_t)(end - p) < l || l != chrNames[k].size() || memcmp(p, chrNames[k].data(), (size_t)l) != 0) { same = false; break; }
			p += l;
			p += (8 - (uintptr_t)(p - (const char *)map) % 8) % 8;
		}
		if (!same || end - p < 8) break;
		int64_t ncomp;
		memcpy(&ncomp, p, 8);
		p += 8;
		t.n_comp = (int32_t)ncomp;
		t.chr_names = chrNames;
		if (!(get_vec(p, end, t.comp_off) && get_vec(p, end, t.comp_chr) && get_vec(p, end, t.comp_start) && get_vec(p, end, t.comp_end) &&
		      get_vec(p, end, t.comp_base_gene) && get_vec(p, end, t.comp_used_gene) && get_vec(p, end, t.se_start) && get_vec(p, end, t.se_end) &&
		      get_vec(p, end, t.se_left_type) && get_vec(p, end, t.se_right_type) && get_vec(p, end, t.se_left_strand) &&
		      get_vec(p, end, t.se_right_strand) && get_vec(p, end, t.se_left_cls) && get_vec(p, end, t.se_right_cls) &&
		      get_vec(p, end, t.se_gene_id) && get_vec(p, end, t.next_off) && get_vec(p, end, t.next_idx) && get_vec(p, end, t.prev_off) &&
		      get_vec(p, end, t.prev_idx)))
			break;
		// shape checks: a damaged file must not reach the kernels
		size_t nse = t.se_start.size();
		if (t.n_comp < 0 || t.comp_off.size() != (size_t)t.n_comp + 1 || t.comp_chr.size() != (size_t)t.n_comp || t.se_end.size() != nse ||
		    t.next_off.size() != nse + 1 || t.prev_off.size() != nse + 1 || (nse && (size_t)t.comp_off.back() != nse) ||
		    (nse && ((size_t)t.next_off.back() != t.next_idx.size() || (size_t)t.prev_off.back() != t.prev_idx.size())))
			break;
		ok = true;
	} while (false);
	munmap(map, (size_t)st.st_size);
	if (ok) g = std::move(t);
	return ok;
}

} // namespace psc
