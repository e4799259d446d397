// psc_internal.h -- device-resident data layout shared by the kernels and the C-ABI host code.
#pragma once
#include <stdint.h>

#include "psc_device.cuh"

namespace pscdev {

// ---- graph (uploaded once; SoA, CSR) ----
struct GraphDev {
	int n_comp, n_se;
	const int *comp_off, *comp_chr, *comp_base, *comp_used;
	const int *comp_max_end;  // largest subexon end of the component
	const int *comp_prev_end; // comp_max_end of the previous component on the same chromosome, or INT_MIN
	const int *se_start, *se_end;
	const int *se_pmax_end; // running maximum of se_end inside the component (== se_end for disjoint subexons)
	const int8_t *lt, *rt, *ls, *rs;
	const double *lcls, *rcls;
	const int *gene;
	const int *next_off, *next_idx, *prev_off, *prev_idx;
};

// ---- one sample: reads (inputs), per-read scratch, and the constraint tables ----
// Rows of group (sample, comp) live at [lo, lo + count) of the per-read arrays, where
// [lo, hi) is the group's contiguous read range: a group can never have more constraints or
// mate pairs than reads, so no compaction or second pass is needed.
struct SampleDev {
	long long n_reads, n_segs;
	int n_chr;
	const long long *chr_off;
	const int *pos, *mate_pos;
	const uint8_t *flags, *nh;
	const uint64_t *name_key;
	const uint32_t *seg_off;
	const int *seg_a, *seg_b;
	int read_len, frag_stdev;

	// group ranges + table sizes, indexed by component
	int *grp_lo, *grp_hi;    // [n_comp]
	long long *bits_off;     // [n_comp+1] offset (in words) of the group's bit rows
	int *n_con, *n_pair;     // [n_comp]

	// per-read scratch
	int4 *cand_meta;     // first (-1: rejected), last, maxReadLen, nh | uniq<<8 | mate kind<<16
	int *cand_cid;       // constraint id of the read
	uint64_t *cand_bits; // [bits_off[n_comp]]
	int *sortbuf;        // [4 * n_reads + 64]: 4R ints of scratch per group (hash tables, NH = 3 sort order)
	uint64_t *pair_key;  // [n_reads]
	int *pair_uniq_in;   // [n_reads]

	// constraint table (struct _constraint, Constraints.hpp:14-26)
	uint64_t *con_bits; // [bits_off[n_comp]]
	int *con_first, *con_last, *con_support, *con_uniq, *con_maxlen;
	int4 *con_nhcnt; // reads per NH value 1..4 (weight = sum of 10/NH, Q1)
	double *con_weight, *con_norm;
	// mate pairs (struct _matePairConstraint, Constraints.hpp:28-39)
	int *pair_i, *pair_j, *pair_support, *pair_uniq, *pair_ec;
	double *pair_norm;
};

} // namespace pscdev
