"""Host logic (BAM reader/writer, graph flattening, sharding, Q8, GTF writer) and the kernel LOGIC
compiled for the CPU by tests/shim (one-lane serial build of the same .cu sources).  These are
debugging / host-logic tests: they say nothing about the CUDA path, which the -m gpu tests cover."""
import os

import pytest

from conftest import ROOT, gtf_body, run, stage_fixture
from psiclass_b200 import testkit

CLASSES_SHIM = os.path.join(ROOT, "tests", "shim", "_build", "classes_shim")


@pytest.mark.parametrize("cfg", [
    dict(seed=41, n_genes=50, n_samples=3),
    dict(seed=42, n_genes=30, n_samples=4, mean_pairs=200.0, exon_mu=2.2, iso_lambda=4.0, p_alt=0.3, p_ir=0.1),
    dict(seed=43, n_genes=40, n_samples=2, single_end=1),
    dict(seed=44, n_genes=40, n_samples=5, p_paralog=0.15, stranded=1, name_style=2, p_noise=0.1, p_overlap=0.15),
    # ~146 subexons per component: three 64-bit words per bit vector, DP path, int32 path-count overflow (Q4)
    dict(seed=20, n_genes=4, n_samples=3, exon_mu=4.8, exon_sigma=0.1, exon_min=110, exon_max=120, iso_lambda=2.0, p_alt=0.3, mean_pairs=150.0),
    # NH = 3: inexact weights, summed in std::sort's order of equal constraints (Q9)
    dict(seed=31, n_genes=60, n_samples=3, mean_pairs=300.0, p_nh3=0.2, iso_lambda=3.0, p_alt=0.3),
    # ... in groups of several 100 000 reads: the capacity-triggered partial coalescing (Constraints.cpp:378-390)
    dict(seed=33, n_genes=3, n_samples=2, mean_pairs=60000.0, p_nh3=0.15, iso_lambda=3.0, p_alt=0.3),
])
def test_kernel_logic_matches_oracle_on_cpu_shim(cfg):
    wl = testkit.Workload.synth(threads=4, **cfg)
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    assert rep.passed, rep.summary()
    assert rep.tables_checked == wl.n_comp * wl.n_samples and rep.val_close == 0


def test_dp_sample_loop_branches_on_cpu_shim():
    """--maxDpConstraintSize -1 on a component of 1587 subexons and 13 samples: DownsampleConstraintsFrom
    (n >= 1500), iterBound = 10 after ten samples with > 1000 candidates, and the skip of such samples
    (TranscriptDecider.cpp:3196-3235)."""
    wl = testkit.Workload.synth(seed=1003, n_genes=1, n_samples=13, dense=1, exon_min=460, exon_max=470, mean_pairs=12.0, threads=4)
    rep = wl.check(testkit.SHIM_PATH, coop=1, max_dp=-1)
    assert rep.passed, rep.summary()
    assert rep.n_downsampled == 13 and rep.n_iter10 >= 1 and rep.n_skipped >= 1


def test_primary_paralog_stranded_many_samples_on_cpu_shim():
    """BASELINE configs[4] in miniature: paralog pairs (NH = 2 primary + secondary), --primaryParalog, rf-stranded
    flags, 13 samples so that the iterBound = 10 rule fires on the heavy DP components."""
    wl = testkit.Workload.synth(seed=1004, n_genes=5, n_samples=13, p_paralog=0.2, stranded=1, mean_pairs=1500.0, exon_mu=3.6,
                                exon_sigma=0.3, iso_lambda=10.0, p_alt=0.5, p_ir=0.1, threads=4)
    rep = wl.check(testkit.SHIM_PATH, coop=1, primary_paralog=1)
    assert rep.passed, rep.summary()
    assert rep.n_iter10 >= 1 and rep.n_dp >= 3


def test_std_sort_order_restatement_equals_libstdcxx(tmp_path):
    """psiclass_b200/csrc/std_sort_order.h makes the moves of libstdc++'s std::sort (needed for the NH = 3 sums)."""
    exe = str(tmp_path / "std_sort_check")
    run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "std_sort_check.cpp")])
    out = run([exe, "6000"])
    assert "mismatches 0" in out, out


def test_two_pass_schedule_keeps_component_major_output(monkeypatch):
    """On a GPU the enumerated components' per-unit kernels run on a second stream under the DP kernel and
    the two result lists are merged back into component order; force that schedule in the serial shim."""
    monkeypatch.setenv("PSC_FORCE_OVERLAP", "1")
    wl = testkit.Workload.synth(seed=42, n_genes=30, n_samples=4, mean_pairs=200.0, exon_mu=2.2, iso_lambda=4.0, p_alt=0.3, p_ir=0.1, threads=4)
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    assert rep.passed, rep.summary()
    monkeypatch.setenv("PSC_MERGE_THREADS", "3")  # the host-side merge of the two lists, multi-threaded
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    assert rep.passed, rep.summary()


@pytest.mark.parametrize("force_overlap", [False, True])
def test_chunked_scratch_budget_gives_the_same_answer(monkeypatch, force_overlap):
    """A 1 MB scratch budget cuts the DP launches and the per-unit passes into many chunks of components."""
    monkeypatch.setenv("PSC_MEM_BUDGET_MB", "1")
    if force_overlap:
        monkeypatch.setenv("PSC_FORCE_OVERLAP", "1")
    wl = testkit.Workload.synth(seed=42, n_genes=30, n_samples=4, mean_pairs=200.0, exon_mu=2.2, iso_lambda=4.0, p_alt=0.3, p_ir=0.1, threads=4)
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    assert rep.passed, rep.summary()


def test_device_pow_algorithm_equals_host_libm(tmp_path):
    """ComputeScore's pow(x, 0.25) is the path's only transcendental; the device evaluates the host libm's own
    algorithm (psiclass_b200/csrc/pow_quarter.h).  Compiled for the host it must equal pow() on every input."""
    exe = str(tmp_path / "pow_check")
    run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "pow_check.cpp")])
    out = run([exe, "20000000"])
    assert "mismatches 0" in out, out


def test_unlimited_dp_constraints_and_paralog_flag():
    wl = testkit.Workload.synth(seed=45, n_genes=25, n_samples=3, mean_pairs=150.0, exon_mu=2.0, iso_lambda=3.0, p_alt=0.3, threads=4)
    assert wl.check(testkit.SHIM_PATH, coop=1, max_dp=-1).passed
    assert wl.check(testkit.SHIM_PATH, coop=1, primary_paralog=1).passed


def test_bam_roundtrip_equals_direct_records(tmp_path):
    """Records pushed straight through the filter == records decoded from the BAM the writer made."""
    wl = testkit.Workload.synth(seed=46, n_genes=30, n_samples=2, name_style=2, p_noise=0.15, threads=2)
    d = str(tmp_path)
    wl.write_files(d)
    w2 = testkit.Workload.load(d)
    assert w2.n_comp == wl.n_comp
    import ctypes as C
    for s in range(wl.n_samples):
        a, b = wl.reads(s), w2.reads(s)
        assert a.n_reads == b.n_reads and a.n_segs == b.n_segs and a.read_len == b.read_len and a.frag_stdev == b.frag_stdev
        n, ns = a.n_reads, a.n_segs
        for fa, fb, k in ((a.pos, b.pos, n), (a.mate_pos, b.mate_pos, n), (a.flags, b.flags, n), (a.nh, b.nh, n),
                          (a.name_key, b.name_key, n), (a.seg_off, b.seg_off, n + 1), (a.seg_a, b.seg_a, ns), (a.seg_b, b.seg_b, ns)):
            assert fa[:k] == fb[:k]
        assert testkit.lib().tk_total_read_cnt(wl.h, s) == testkit.lib().tk_total_read_cnt(w2.h, s)


FIXTURES = {"synth_small": [], "dp_small": [], "nh3_small": [], "example": [], "dense_down": ["--maxDpConstraintSize", "-1"]}


@pytest.mark.parametrize("name", sorted(FIXTURES))
@pytest.mark.parametrize("shards", [1, 3])
def test_front_end_reproduces_reference_gtfs(name, shards, tmp_path):
    """argv handling, single-pass BAM decode, component sharding, Q8 and the GTF writer: the drop-in
    `classes` (linked against the shim here) must print the reference's GTF bodies byte for byte."""
    d = stage_fixture(name, tmp_path)
    env = dict(os.environ, PSC_GPUS=str(shards))
    out = run([CLASSES_SHIM, "--lb", "bamlist", "-s", "subexon_combined.out", "-o", "new", "-p", "3"] + FIXTURES[name], cwd=d, env=env)
    n = len(open(os.path.join(d, "bamlist")).read().split())
    for s in range(n):
        assert gtf_body(os.path.join(d, "new_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s))
    golden = [ln for ln in open(os.path.join(d, "ref_dump.txt")).read().splitlines() if "atCnt=" in ln]
    assert [ln for ln in out.splitlines() if "atCnt=" in ln] == golden


def test_front_end_error_paths(tmp_path):
    import subprocess
    r = subprocess.run([CLASSES_SHIM], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "Required:" in r.stdout            # classes.cpp:102-106
    r = subprocess.run([CLASSES_SHIM, "-b", "x.bam"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Cannot find combined subexon file." in r.stdout   # classes.cpp:210-214
    r = subprocess.run([CLASSES_SHIM, "--nonsense"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1                                        # classes.cpp:204-208
    d = stage_fixture("synth_small", tmp_path)
    r = subprocess.run([CLASSES_SHIM, "-s", "subexon_combined.out"], cwd=d, stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Must use -b option" in r.stdout   # classes.cpp:215-219


def test_empty_and_ragged_inputs():
    # a sample with (almost) no reads, genes without reads, single-subexon genes
    wl = testkit.Workload.synth(seed=47, n_genes=30, n_samples=3, mean_pairs=0.3, exon_min=1, exon_max=2, threads=2)
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    assert rep.passed, rep.summary()


def test_graph_cache_round_trip_and_invalidation(tmp_path):
    """the parsed subexon graph as a binary file (<subexon file>.pscgraph, PSC_GRAPH_CACHE=1).  A valid
    cache gives the same GTFs as parsing; a truncated one, one made for another classifier threshold or for another
    source file is ignored."""
    d = stage_fixture("dp_small", tmp_path)
    env = dict(os.environ, PSC_GRAPH_CACHE="1")
    cache = os.path.join(d, "subexon_combined.out.pscgraph")
    n = len(open(os.path.join(d, "bamlist")).read().split())

    def classes(prefix, extra=(), e=env):
        run([CLASSES_SHIM, "--lb", "bamlist", "-s", "subexon_combined.out", "-o", prefix, "-p", "2"] + list(extra), cwd=d, env=e)
        return [gtf_body(os.path.join(d, "%s_sample_%d.gtf" % (prefix, s))) for s in range(n)]

    ref = [gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s)) for s in range(n)]
    assert classes("a") == ref and os.path.exists(cache)          # parsed, cache written
    stamp = os.path.getmtime(cache)
    assert classes("b") == ref and os.path.getmtime(cache) == stamp  # read back from the cache, not rewritten
    # another threshold: the cache is for -c 0.05, so it is ignored (and replaced)
    plain = classes("c", ["-c", "0.5"], dict(os.environ))
    assert classes("d", ["-c", "0.5"]) == plain
    # truncated cache: ignored
    with open(cache, "r+b") as f:
        f.truncate(200)
    assert classes("e") == ref
    # source changed after the cache was made: ignored
    assert classes("f") == ref
    stamp = os.path.getmtime(cache)
    os.utime(os.path.join(d, "subexon_combined.out"), (1, 1))
    assert classes("g") == ref and os.path.getmtime(cache) != stamp   # reparsed and rewritten for the new stamp


@pytest.mark.parametrize("cfg", [
    dict(seed=41, n_genes=200, n_samples=1, p_overlap=0.2),
    dict(seed=44, n_genes=150, n_samples=1, p_paralog=0.15, exon_mu=2.4, iso_lambda=4.0, p_alt=0.4, p_ir=0.1),
    dict(seed=1003, n_genes=3, n_samples=1, dense=1, exon_min=80, exon_max=100),
])
def test_subexon_file_writer_and_cache_round_trip(cfg):
    """the graph written in the combined-subexon text format (CombineSubexons.cpp:1617-1684) and as the
    binary cache reads back to the same components, gene ids, classifiers and edges."""
    wl = testkit.Workload.synth(threads=2, mean_pairs=1.0, **cfg)
    assert wl.graph_roundtrip() == 0


@pytest.mark.parametrize("name", ["synth_small", "dp_small", "example"])
def test_subexon_file_round_trip_on_fixtures(name, tmp_path):
    wl = testkit.Workload.load(stage_fixture(name, tmp_path))
    assert wl.graph_roundtrip() == 0
