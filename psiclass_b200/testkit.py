"""ctypes access to build/libpsc_testkit.so (TEST / BENCH INFRASTRUCTURE): synthetic workloads,
file loading, the CPU oracle and the engine-vs-oracle comparison.  Not used by the product."""
import ctypes as C
import os

from . import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TK_PATH = os.path.join(ROOT, "build", "libpsc_testkit.so")
SHIM_PATH = os.path.join(ROOT, "tests", "shim", "_build", "libpsc_shim.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class SynthCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_chr", C.c_int32), ("n_genes", C.c_int32), ("n_samples", C.c_int32),
                ("mean_pairs", C.c_double), ("exon_mu", C.c_double), ("exon_sigma", C.c_double),
                ("exon_min", C.c_int32), ("exon_max", C.c_int32), ("iso_lambda", C.c_double),
                ("p_skip", C.c_double), ("p_alt", C.c_double), ("p_ir", C.c_double), ("p_noise", C.c_double),
                ("p_overlap", C.c_double), ("p_paralog", C.c_double), ("name_style", C.c_int32),
                ("stranded", C.c_int32), ("dense", C.c_int32), ("graph_noise", C.c_int32),
                ("single_end", C.c_int32), ("threads", C.c_int32), ("p_nh3", C.c_double), ("frag_sd", C.c_double)]


class CheckOpts(C.Structure):
    _fields_ = [("maxDp", C.c_int), ("coop", C.c_int), ("primaryParalog", C.c_int), ("constraintsOnly", C.c_int),
                ("verbose", C.c_int), ("device", C.c_int)]


class CheckReport(C.Structure):
    _fields_ = [("passed", C.c_int), ("n_comp", C.c_int), ("n_dp", C.c_int),
                ("tables_checked", C.c_longlong), ("tables_bad", C.c_longlong), ("stats_bad", C.c_longlong),
                ("tx_engine", C.c_longlong), ("tx_oracle", C.c_longlong), ("tx_bad", C.c_longlong),
                ("val_exact", C.c_longlong), ("val_close", C.c_longlong), ("val_bad", C.c_longlong),
                ("launches", C.c_longlong), ("constraints_ms", C.c_double), ("solve_ms", C.c_double),
                ("n_downsampled", C.c_longlong), ("n_iter10", C.c_longlong), ("n_skipped", C.c_longlong)]

    def summary(self):
        return ("pass=%d comps=%d dp=%d tables %d/%d bad, stats bad %d, tx engine/oracle %d/%d bad %d, "
                "values exact/close/off %d/%d/%d" % (
                    self.passed, self.n_comp, self.n_dp, self.tables_bad, self.tables_checked, self.stats_bad,
                    self.tx_engine, self.tx_oracle, self.tx_bad, self.val_exact, self.val_close, self.val_bad))


_tk = None


def lib():
    global _tk
    if _tk is None:
        if not os.path.exists(TK_PATH):
            raise RuntimeError("%s not built (run __graft_entry__.build())" % TK_PATH)
        t = C.CDLL(TK_PATH)
        t.tk_synth.restype = C.c_void_p
        t.tk_synth.argtypes = [C.POINTER(SynthCfg)]
        t.tk_load.restype = C.c_void_p
        t.tk_load.argtypes = [C.c_char_p]
        t.tk_error.restype = C.c_char_p
        t.tk_error.argtypes = [C.c_void_p]
        t.tk_free.argtypes = [C.c_void_p]
        t.tk_n_samples.argtypes = [C.c_void_p]
        t.tk_n_comp.argtypes = [C.c_void_p]
        t.tk_n_reads.restype = C.c_longlong
        t.tk_n_reads.argtypes = [C.c_void_p, C.c_int]
        t.tk_get_graph.argtypes = [C.c_void_p, C.POINTER(abi.psc_graph)]
        t.tk_get_reads.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.psc_reads)]
        t.tk_get_params.argtypes = [C.c_void_p, C.POINTER(abi.psc_params)]
        t.tk_write_files.argtypes = [C.c_void_p, C.c_char_p]
        t.tk_check.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(CheckOpts), C.POINTER(CheckReport)]
        t.tk_total_read_cnt.restype = C.c_double
        t.tk_total_read_cnt.argtypes = [C.c_void_p, C.c_int]
        t.tk_default_cfg.argtypes = [C.POINTER(SynthCfg)]
        t.tk_graph_roundtrip.argtypes = [C.c_void_p]
        t.tk_write_subexon_file.argtypes = [C.c_void_p, C.c_char_p]
        _tk = t
    return _tk


class Workload:
    def __init__(self, handle):
        self.h = handle
        err = lib().tk_error(handle)
        if err:
            raise RuntimeError(err.decode())

    @classmethod
    def synth(cls, **kw):
        c = SynthCfg()
        lib().tk_default_cfg(C.byref(c))
        for k, v in kw.items():
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
        return cls(lib().tk_synth(C.byref(c)))

    @classmethod
    def load(cls, directory):
        return cls(lib().tk_load(directory.encode()))

    @property
    def n_samples(self):
        return lib().tk_n_samples(self.h)

    @property
    def n_comp(self):
        return lib().tk_n_comp(self.h)

    def n_reads(self, s):
        return lib().tk_n_reads(self.h, s)

    def graph(self):
        g = abi.psc_graph()
        lib().tk_get_graph(self.h, C.byref(g))
        return g

    def reads(self, s):
        r = abi.psc_reads()
        lib().tk_get_reads(self.h, s, C.byref(r))
        return r

    def params(self):
        p = abi.psc_params()
        lib().tk_get_params(self.h, C.byref(p))
        return p

    def write_files(self, directory):
        rc = lib().tk_write_files(self.h, directory.encode())
        if rc:
            raise RuntimeError("tk_write_files failed: %d" % rc)

    def check(self, engine_lib, coop=32, max_dp=7, primary_paralog=0, constraints_only=0, verbose=0, device=0):
        o = CheckOpts(max_dp, coop, primary_paralog, constraints_only, verbose, device)
        r = CheckReport()
        rc = lib().tk_check(self.h, engine_lib.encode(), C.byref(o), C.byref(r))
        if rc >= 100:
            raise RuntimeError(lib().tk_error(self.h).decode())
        return r

    def write_subexon_file(self, path):
        if lib().tk_write_subexon_file(self.h, path.encode()):
            raise RuntimeError("cannot write " + path)

    def graph_roundtrip(self):
        """0 when writing the graph as text / as the binary cache and reading it back reproduces it"""
        return lib().tk_graph_roundtrip(self.h)

    def close(self):
        if self.h:
            lib().tk_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
