// Host-side subexon graph of stage 3, flattened to CSR per gene component.
//
// Replaces the reference's SubexonGraph (/root/reference/SubexonGraph.hpp:33-346,
// SubexonGraph.cpp:1-241): parsing of the combined subexon file, classifier filter,
// coordinate -> index edge resolution, gene-interval (component) discovery and the
// sequential gene-id assignment.  The result is the structure-of-arrays layout the CUDA
// kernels read (include/psc.h: psc_graph).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace psc {

struct Graph {
	std::vector<std::string> chr_names; // from BAM #0 (classes.cpp:269)

	// ---- components (gene intervals), in genome order ----
	int32_t n_comp = 0;
	std::vector<int32_t> comp_off;       // [n_comp+1] offset into the per-subexon arrays
	std::vector<int32_t> comp_chr;       // chrId
	std::vector<int32_t> comp_start;     // _geneInterval.start (0-based, may be adjusted)
	std::vector<int32_t> comp_end;       // _geneInterval.end
	std::vector<int32_t> comp_base_gene; // baseGeneId  (reference keeps it in subexons[0].lcCnt)
	std::vector<int32_t> comp_used_gene; // usedGeneId  (reference keeps it in subexons[0].rcCnt)

	// ---- subexons (global order; adjacency holds component-LOCAL indices) ----
	std::vector<int32_t> se_start, se_end; // 0-based inclusive
	std::vector<int8_t> se_left_type, se_right_type;     // 0 soft, 1 hard start, 2 hard end
	std::vector<int8_t> se_left_strand, se_right_strand; // -1, 0, 1
	std::vector<double> se_left_cls, se_right_cls;       // classifiers
	std::vector<int32_t> se_gene_id;                     // -2 = shared between genes
	std::vector<int32_t> next_off, next_idx;             // [n_se+1], local successor indices
	std::vector<int32_t> prev_off, prev_idx;

	int32_t n_se() const { return (int32_t)se_start.size(); }
};

// Parse `path` and build the graph.  `chr_names` gives the chromosome ids (BAM #0 header).
// Unknown chromosome -> false with err = "Unknown genome name: X" (alignments.hpp:290-299).
bool load_subexon_graph(const std::string &path, const std::vector<std::string> &chr_names,
                        double classifier_threshold, Graph &g, std::string &err);

// Writes the graph in the combined-subexon text format (/root/reference/CombineSubexons.cpp:1617-1684: `chr start end
// leftType rightType leftStrand rightStrand -1 -1 -1 leftClassifier rightClassifier prevCnt prev... nextCnt next...`,
// 1-based coordinates, prev = end coordinate of a predecessor, next = start coordinate of a successor, classifiers
// with 17 significant digits).  What is written is the graph as stage 3 uses it (after the classifier filter and the
// fix-ups of the reader); reading the file back gives the same graph.
bool write_subexon_file(const Graph &g, const std::string &path);

// Binary form of a parsed graph: the arrays above, little-endian, behind a header that pins the source
// file (size, mtime), the classifier threshold and the chromosome names it was built for.  `load_graph_cache` maps the
// file and copies the arrays out; it returns false (and leaves `g` untouched) when the file is missing, truncated or
// was made from a different source / threshold / header.  `classes` uses `<subexon file>.pscgraph` when the
// environment has PSC_GRAPH_CACHE=1: read it if valid, else parse the text and write it.
bool save_graph_cache(const std::string &cache_path, const std::string &source_path, double classifier_threshold, const Graph &g);
bool load_graph_cache(const std::string &cache_path, const std::string &source_path, const std::vector<std::string> &chr_names,
                      double classifier_threshold, Graph &g);

} // namespace psc
