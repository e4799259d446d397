/* psc.h -- C ABI of the B200-native PsiCLASS stage-3 engine (libpsc.so).
 *
 * The reference (splicebox/PsiCLASS v1.0.3) exposes no FFI: its boundary is the `classes`
 * process (classes.cpp:97-508) and, inside it, three C++ interfaces that this ABI replaces:
 *
 *   SubexonGraph::ExtractSubexons()            SubexonGraph.cpp:114-149   -> psc_set_graph
 *   Constraints::BuildConstraints()            Constraints.cpp:240-438    -> psc_set_reads +
 *                                                                            psc_build_constraints
 *   TranscriptDecider::Solve()                 TranscriptDecider.cpp:2874 -> psc_solve
 *   MultiThreadOutputTranscript::Add()         TranscriptDecider.hpp:178  -> psc_get_transcripts
 *
 * Plain C types only; every buffer is caller-owned host memory; every call returns 0 on success
 * and a non-zero code otherwise (psc_last_error() gives the message; the `classes` front end
 * prints it and exit(1)s like the reference does on its fatal paths, classes.cpp:210-219).
 * All calls on one context must come from one host thread.  There is no CPU fallback: creating
 * a context without a CUDA device fails.
 */
#ifndef PSC_H
#define PSC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psc_ctx psc_ctx;

/* Run parameters (classes.cpp:111-119, TranscriptDecider.hpp:480-492).  The library statistics
 * come from BAM #0 for every sample (TranscriptDecider keeps alignmentFiles[0]; classes.cpp:287). */
typedef struct psc_params {
	double fpkm_fraction;        /* -f, default 0.05 */
	double classifier_threshold; /* -c, default 0.05 */
	double min_read_depth;       /* -d, default 2.5  */
	int32_t max_dp_constraint_size; /* --maxDpConstraintSize, default 7; <=0: unlimited */
	int32_t primary_as_unique;      /* --primaryParalog */
	int32_t n_samples;
	int32_t read_len, frag_len, frag_stdev, mate_paired; /* Alignments of BAM #0 */
} psc_params;

/* The subexon graph flattened to CSR per gene component (struct _subexon, SubexonGraph.hpp:7-24).
 * Adjacency lists hold component-local subexon indices. */
typedef struct psc_graph {
	int32_t n_comp, n_se;
	const int32_t *comp_off;       /* [n_comp+1] first subexon of every component          */
	const int32_t *comp_chr;       /* [n_comp]   chromosome id                              */
	const int32_t *comp_base_gene; /* [n_comp]   baseGeneId (reference: subexons[0].lcCnt)  */
	const int32_t *comp_used_gene; /* [n_comp]   usedGeneId (reference: subexons[0].rcCnt)  */
	const int32_t *se_start;       /* [n_se] 0-based inclusive */
	const int32_t *se_end;
	const int8_t *se_left_type, *se_right_type;     /* 0 soft, 1 hard start, 2 hard end */
	const int8_t *se_left_strand, *se_right_strand; /* -1, 0, +1 */
	const double *se_left_cls, *se_right_cls;       /* classifiers */
	const int32_t *se_gene_id;                      /* -2: shared by several genes */
	const int32_t *next_off, *next_idx;             /* [n_se+1], [next_off[n_se]] */
	const int32_t *prev_off, *prev_idx;
} psc_graph;

/* flag bits of psc_reads.flags */
#define PSC_RF_PRIMARY 1
#define PSC_RF_CLIP_HEAD 2
#define PSC_RF_CLIP_TAIL 4
#define PSC_RF_FIRST_MATE 8
#define PSC_RF_UNIQUE 16
#define PSC_RF_MATE_SAME_CHR 32

/* One sample's alignments: exactly the records for which Alignments::Next() returns 1
 * (alignments.hpp:146-277), in file (coordinate) order, structure of arrays. */
typedef struct psc_reads {
	int64_t n_reads, n_segs;
	int32_t n_chr;
	const int64_t *chr_off;   /* [n_chr+1] first read of every chromosome */
	const int32_t *pos;       /* [n_reads] segments[0].a */
	const int32_t *mate_pos;  /* [n_reads] core.mpos */
	const uint8_t *flags;     /* [n_reads] PSC_RF_* */
	const uint8_t *nh;        /* [n_reads] NH tag (1 if absent) */
	const uint64_t *name_key; /* [n_reads] 64-bit hash of the read id used for mate lookup */
	const uint32_t *seg_off;  /* [n_reads+1] */
	const int32_t *seg_a, *seg_b; /* [n_segs] aligned blocks, 0-based inclusive */
	int32_t read_len, frag_stdev; /* per-sample: Constraints.cpp:203-204 and :405 */
} psc_reads;

/* One (component, sample) constraint table, for parity checks (struct _constraint and
 * _matePairConstraint, Constraints.hpp:14-39). */
typedef struct psc_constraint_table {
	int32_t n_constraints, n_pairs, n_words; /* n_words = ceil(seCnt/64) */
	uint64_t *bits;     /* [n_constraints * n_words] */
	int32_t *first, *last, *support, *uniq_support, *max_read_len;
	double *weight, *norm_abund;
	int32_t *pair_i, *pair_j, *pair_support, *pair_uniq, *pair_effective_count;
	double *pair_norm_abund;
} psc_constraint_table;

/* Output transcripts in the reference's queue order: component-major, then sample, then the
 * order of predTranscripts (TranscriptDecider.cpp:3432-3438).  fpkm is the value before
 * ComputeFPKMTPM (TranscriptDecider.hpp:190-208). */
typedef struct psc_transcripts {
	int64_t n_transcripts, n_exons;
	int32_t *comp, *sample, *chr, *gene_id;
	int8_t *strand;        /* '+', '-', '.' */
	int64_t *exon_off;     /* [n_transcripts+1] */
	int32_t *exon_start, *exon_end; /* 1-based inclusive */
	double *fpkm, *cov;
} psc_transcripts;

/* per-component statistics (the `atCnt=` log line, TranscriptDecider.cpp:3111) */
typedef struct psc_comp_stats {
	int32_t *at_cnt, *use_dp, *n_candidates; /* [n_comp] */
} psc_comp_stats;

int psc_create(psc_ctx **out, int device);
void psc_destroy(psc_ctx *ctx);
const char *psc_last_error(const psc_ctx *ctx);

int psc_set_params(psc_ctx *ctx, const psc_params *p);
int psc_set_graph(psc_ctx *ctx, const psc_graph *g);
/* Copies the sample's reads host->device: the copies are only ENQUEUED on the context's stream, so the caller's arrays
 * must stay valid and unchanged until the next synchronising call on the context (psc_build_constraints).  Pinned arrays
 * upload at full PCIe speed.  Replaces the per-sample Alignments stream of classes.cpp:264-283. */
int psc_set_reads(psc_ctx *ctx, int32_t sample, const psc_reads *r);

/* Alignment -> constraint tables for every (sample, component). */
int psc_build_constraints(psc_ctx *ctx);
/* counts[sample * n_comp + comp] */
int psc_constraint_counts(psc_ctx *ctx, int32_t *n_constraints, int32_t *n_pairs);
/* Copies one table into caller buffers sized from psc_constraint_counts. */
int psc_get_constraint_table(psc_ctx *ctx, int32_t sample, int32_t comp, psc_constraint_table *out);

/* Candidate generation (enumeration / DP), per-sample set-cover picking, refinement, EM and
 * transcript emission for every component. */
int psc_solve(psc_ctx *ctx);
int psc_transcript_counts(psc_ctx *ctx, int64_t *n_transcripts, int64_t *n_exons);
int psc_get_transcripts(psc_ctx *ctx, psc_transcripts *out);
int psc_get_comp_stats(psc_ctx *ctx, psc_comp_stats *out);

/* Device timing of the last psc_build_constraints / psc_solve (CUDA events on the context's
 * stream), kernel launch count, and algorithmic bytes (DESIGN.md "roofline"). */
typedef struct psc_timing {
	double constraints_ms, solve_ms;
	int64_t kernel_launches;
	int64_t constraint_bytes; /* algorithmic bytes of the constraint kernel, last call */
	/* per-kernel device time (CUDA events around every launch, summed) and launch counts of the last
	 * psc_build_constraints + psc_solve: 0 group_ranges, 1 build_constraints, 2 prologue,
	 * 3 enumerate, 4 dp_units, 5 units phase 1 (pick round 1), 6 emit, 7 dp_merge,
	 * 8 units phase 2 (pick round 2, refine, EM), 9 units phase 3 (rescue), 10 dp2_units (static-program DP),
	 * 11 dp2_units launched one thread per unit (experiment switch PSC_DP_T1_FRAC, off by default) */
	double kernel_ms[12];
	int64_t kernel_count[12];
	int64_t kernel_bytes[12]; /* algorithmic bytes attributed to the kernel, 0 if not modelled */
} psc_timing;
int psc_get_timing(psc_ctx *ctx, psc_timing *out);

/* Cooperation width of the solve kernels: 32 (one warp per unit, default) or 1 (one thread per
 * unit; used by the tests to cross-check the warp-cooperative code paths). */
int psc_set_coop_width(psc_ctx *ctx, int32_t width);

/* Diagnostic: y[i] = the device's pow(x[i], 0.25) as ComputeScore (TranscriptDecider.hpp:394-401) evaluates it,
 * x in (0.1, 1].  The tests compare it with the host libm, whose rounding the engine reproduces bit for bit. */
int psc_debug_pow_quarter(psc_ctx *ctx, const double *x, double *y, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* PSC_H */
