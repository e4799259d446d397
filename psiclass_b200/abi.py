"""ctypes mirror of include/psc.h -- the C ABI of libpsc.so (the CUDA engine).

This module is plumbing for tests/ and bench.py: the product is the shared library and the
`classes` front end in psiclass_b200/bin.  There is no CPU fallback: `load()` raises if the
library has not been built by nvcc, and `psc_create` fails without a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSC_LIBPSC") or os.path.join(HERE, "libpsc.so")  # PSC_LIBPSC: build variants (tools/gpu)

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
i8p = C.POINTER(C.c_int8)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class psc_params(C.Structure):
    _fields_ = [("fpkm_fraction", C.c_double), ("classifier_threshold", C.c_double),
                ("min_read_depth", C.c_double), ("max_dp_constraint_size", C.c_int32),
                ("primary_as_unique", C.c_int32), ("n_samples", C.c_int32),
                ("read_len", C.c_int32), ("frag_len", C.c_int32), ("frag_stdev", C.c_int32),
                ("mate_paired", C.c_int32)]


class psc_graph(C.Structure):
    _fields_ = [("n_comp", C.c_int32), ("n_se", C.c_int32),
                ("comp_off", i32p), ("comp_chr", i32p), ("comp_base_gene", i32p), ("comp_used_gene", i32p),
                ("se_start", i32p), ("se_end", i32p),
                ("se_left_type", i8p), ("se_right_type", i8p), ("se_left_strand", i8p), ("se_right_strand", i8p),
                ("se_left_cls", f64p), ("se_right_cls", f64p), ("se_gene_id", i32p),
                ("next_off", i32p), ("next_idx", i32p), ("prev_off", i32p), ("prev_idx", i32p)]


class psc_reads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_segs", C.c_int64), ("n_chr", C.c_int32),
                ("chr_off", i64p), ("pos", i32p), ("mate_pos", i32p), ("flags", u8p), ("nh", u8p),
                ("name_key", u64p), ("seg_off", u32p), ("seg_a", i32p), ("seg_b", i32p),
                ("read_len", C.c_int32), ("frag_stdev", C.c_int32)]


class psc_constraint_table(C.Structure):
    _fields_ = [("n_constraints", C.c_int32), ("n_pairs", C.c_int32), ("n_words", C.c_int32),
                ("bits", u64p), ("first", i32p), ("last", i32p), ("support", i32p), ("uniq_support", i32p),
                ("max_read_len", i32p), ("weight", f64p), ("norm_abund", f64p),
                ("pair_i", i32p), ("pair_j", i32p), ("pair_support", i32p), ("pair_uniq", i32p),
                ("pair_effective_count", i32p), ("pair_norm_abund", f64p)]


class psc_transcripts(C.Structure):
    _fields_ = [("n_transcripts", C.c_int64), ("n_exons", C.c_int64),
                ("comp", i32p), ("sample", i32p), ("chr", i32p), ("gene_id", i32p), ("strand", i8p),
                ("exon_off", i64p), ("exon_start", i32p), ("exon_end", i32p), ("fpkm", f64p), ("cov", f64p)]


class psc_comp_stats(C.Structure):
    _fields_ = [("at_cnt", i32p), ("use_dp", i32p), ("n_candidates", i32p)]


class psc_timing(C.Structure):
    _fields_ = [("constraints_ms", C.c_double), ("solve_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("constraint_bytes", C.c_int64), ("kernel_ms", C.c_double * 12), ("kernel_count", C.c_int64 * 12), ("kernel_bytes", C.c_int64 * 12)]


# every symbol include/psc.h declares
SYMBOLS = ["psc_create", "psc_destroy", "psc_last_error", "psc_set_params", "psc_set_graph", "psc_set_reads",
           "psc_build_constraints", "psc_constraint_counts", "psc_get_constraint_table", "psc_solve",
           "psc_transcript_counts", "psc_get_transcripts", "psc_get_comp_stats", "psc_get_timing",
           "psc_set_coop_width", "psc_debug_pow_quarter"]


def load(path=None):
    """Load libpsc.so and declare the prototypes.  Raises if the CUDA library is not built."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.psc_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.psc_create.restype = C.c_int
    lib.psc_destroy.argtypes = [vp]
    lib.psc_destroy.restype = None
    lib.psc_last_error.argtypes = [vp]
    lib.psc_last_error.restype = C.c_char_p
    lib.psc_set_params.argtypes = [vp, C.POINTER(psc_params)]
    lib.psc_set_graph.argtypes = [vp, C.POINTER(psc_graph)]
    lib.psc_set_reads.argtypes = [vp, C.c_int32, C.POINTER(psc_reads)]
    lib.psc_build_constraints.argtypes = [vp]
    lib.psc_constraint_counts.argtypes = [vp, i32p, i32p]
    lib.psc_get_constraint_table.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(psc_constraint_table)]
    lib.psc_solve.argtypes = [vp]
    lib.psc_transcript_counts.argtypes = [vp, i64p, i64p]
    lib.psc_get_transcripts.argtypes = [vp, C.POINTER(psc_transcripts)]
    lib.psc_get_comp_stats.argtypes = [vp, C.POINTER(psc_comp_stats)]
    lib.psc_get_timing.argtypes = [vp, C.POINTER(psc_timing)]
    lib.psc_set_coop_width.argtypes = [vp, C.c_int32]
    lib.psc_debug_pow_quarter.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64]
    for name in SYMBOLS:
        if name not in ("psc_destroy", "psc_last_error"):
            getattr(lib, name).restype = C.c_int
    return lib


class Engine:
    """Thin RAII wrapper: one context on one CUDA device."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load()
        self.ctx = C.c_void_p()
        rc = self.lib.psc_create(C.byref(self.ctx), device)
        if rc != 0:
            raise RuntimeError("psc_create failed with code %d (no CUDA device? there is no CPU fallback)" % rc)

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.ctx, *args)
        if rc != 0:
            raise RuntimeError("%s: %s" % (name, self.lib.psc_last_error(self.ctx).decode()))

    def timing(self):
        t = psc_timing()
        self.lib.psc_get_timing(self.ctx, C.byref(t))
        return t

    def close(self):
        if self.ctx:
            self.lib.psc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
