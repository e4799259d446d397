"""The CPU oracle (oracle/psc_oracle.cpp) and the host-side parsers against the UNMODIFIED reference:
committed golden fixtures (made by tests/golden/make_golden.sh with oracle/_ref) and, where
oracle/_ref is present, fresh seeded sets diffed live."""
import os

import pytest

from conftest import ROOT, gtf_body, run, stage_fixture

ORACLE = os.path.join(ROOT, "oracle", "_ref", "psc_oracle")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
SYNTH = os.path.join(ROOT, "build", "synth")


def dump_lines(text):
    return [ln for ln in text.splitlines() if "atCnt=" not in ln]


def atcnt_lines(text):
    return [ln for ln in text.splitlines() if "atCnt=" in ln]


# fixture -> extra `classes` arguments it was generated with (tests/golden/make_golden.sh)
FIXTURES = {"synth_small": [], "dp_small": [], "nh3_small": [], "example": [],
            "dense_down": ["--maxDpConstraintSize", "-1"]}


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_oracle_matches_golden(name, tmp_path):
    d = stage_fixture(name, tmp_path)
    out = run([ORACLE, "-s", "subexon_combined.out", "--lb", "bamlist", "-o", "orc", "--dump"] + FIXTURES[name], cwd=d)
    golden = open(os.path.join(d, "ref_dump.txt")).read()
    if name == "dense_down":  # the fixture keeps no constraint / pair rows (size); everything else is compared
        out = "\n".join(ln for ln in out.splitlines() if not ln.startswith(("c ", "m ")))
    # library stats, component list + gene ids, every constraint / mate-pair table, totalReadCnt (Q8)
    assert dump_lines(out) == dump_lines(golden)
    # atCnt / seCnt / useDP / #constraints / #matePairs of every Solve
    assert atcnt_lines(out) == atcnt_lines(golden)
    n = len(open(os.path.join(d, "bamlist")).read().split())
    for s in range(n):
        assert gtf_body(os.path.join(d, "orc_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s))


def test_example_known_answer(tmp_path):
    """the reference's example/ (only where /root/reference is mounted)."""
    ex = "/root/reference/example"
    if not os.path.isdir(ex) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "combine-subexons")):
        pytest.skip("reference example not available")
    d = str(tmp_path)
    with open(os.path.join(d, "bamlist"), "w") as f:
        f.write(ex + "/s1.bam\n" + ex + "/s2.bam\n")
    run([os.path.join(ROOT, "oracle", "_ref", "psiclass"), "--lb", "bamlist", "-o", "ex_", "--stage", "0"], cwd=d)
    sub = os.path.join(d, "subexon", "ex_subexon_combined.out")
    out = run([ORACLE, "-s", sub, "--lb", "bamlist", "-o", "orc", "--dump"], cwd=d)
    lines = out.splitlines()
    assert "S 0 0 22 28" in lines and "T 0 192" in lines and "T 1 126" in lines
    assert "c 1 0 1 18 18 100 180 2.278481012658228 2 0 1" in lines
    assert "m 1 0 2 5 5 2 2.4050632911392404" in lines
    body = gtf_body(os.path.join(d, "orc_sample_0.gtf"))
    assert 'FPKM "240454.413473"; TPM "1000000.000000"; cov "9.233449";' in body[0]
    assert body == gtf_body(os.path.join(d, "ex_sample_0.gtf"))


@pytest.mark.parametrize("args", [
    ["--genes", "60", "--samples", "3", "--seed", "31", "--name-style", "2", "--p-noise", "0.1", "--p-overlap", "0.15"],
    ["--genes", "30", "--samples", "4", "--seed", "32", "--pairs", "200", "--exon-mu", "2.2", "--iso-lambda", "4", "--p-alt", "0.3"],
    ["--genes", "40", "--samples", "2", "--seed", "33", "--single-end", "1"],
    # NH = 3 reads in groups deep enough (> 32768 constraints) for the capacity-triggered partial coalescing
    ["--genes", "3", "--samples", "2", "--seed", "33", "--pairs", "60000", "--p-nh3", "0.15", "--iso-lambda", "3", "--p-alt", "0.3"],
])
def test_oracle_matches_live_reference(args, tmp_path):
    if not os.path.exists(REF_DUMP):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    d = str(tmp_path)
    run([SYNTH, "-o", d] + args)
    ref = run([REF_DUMP, "-s", "subexon_combined.out", "--lb", "bamlist", "-o", "ref"], cwd=d)
    orc = run([ORACLE, "-s", "subexon_combined.out", "--lb", "bamlist", "-o", "orc", "--dump"], cwd=d)
    assert dump_lines(orc) == dump_lines(ref)
    assert atcnt_lines(orc) == atcnt_lines(ref)
    n = len(open(os.path.join(d, "bamlist")).read().split())
    for s in range(n):
        assert gtf_body(os.path.join(d, "orc_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s))


@pytest.mark.parametrize("name", ["synth_small", "dp_small"])
def test_reference_reads_the_written_subexon_file(name, tmp_path):
    """The subexon-file writer (psc::write_subexon_file): the UNMODIFIED reference, run on the file we write from the
    parsed graph, produces the same tables and GTFs as on the original combined subexon file."""
    if not os.path.exists(REF_DUMP):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from psiclass_b200 import testkit
    d = stage_fixture(name, tmp_path)
    testkit.Workload.load(d).write_subexon_file(os.path.join(d, "rewritten.out"))
    out = run([REF_DUMP, "-s", "rewritten.out", "--lb", "bamlist", "-o", "rw"], cwd=d)
    golden = open(os.path.join(d, "ref_dump.txt")).read()
    assert dump_lines(out) == dump_lines(golden)
    assert atcnt_lines(out) == atcnt_lines(golden)
    n = len(open(os.path.join(d, "bamlist")).read().split())
    for s in range(n):
        assert gtf_body(os.path.join(d, "rw_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s))
