"""Stage 4 (vote) and the whole drop-in under the reference's own driver script.

* psiclass_b200/bin/vote-transcripts and `classes --vote` (fused, in memory) against the reference's
  vote-transcripts (Vote.cpp) -- committed golden vote of example/ plus live runs where oracle/_ref exists;
* `psiclass` (the unmodified Perl driver from oracle/_ref) with its stage-3 and stage-4 executables replaced by ours,
  stages 0-4 on the reference's example: vote.gtf and the per-sample GTFs must equal the reference's."""
import os
import shutil
import subprocess

import pytest

from conftest import GOLDEN, ROOT, gtf_body, run, stage_fixture

REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "psiclass_b200", "bin")
SHIM = os.path.join(ROOT, "tests", "shim", "_build")
FIXTURES = {"synth_small": [], "dp_small": [], "nh3_small": [], "example": [], "dense_down": ["--maxDpConstraintSize", "-1"]}


def gtf_list(d, prefix):
    n = len(open(os.path.join(d, "bamlist")).read().split())
    p = os.path.join(d, prefix + "_gtf.list")
    with open(p, "w") as f:
        for s in range(n):
            f.write(os.path.join(d, "%s_sample_%d.gtf" % (prefix, s)) + "\n")
    return p


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_vote_binary_matches_reference(name, tmp_path):
    d = stage_fixture(name, tmp_path)
    lst = gtf_list(d, "ref")
    ours = run([os.path.join(BIN, "vote-transcripts"), "--lg", lst])
    if name == "example":
        assert ours == open(os.path.join(d, "ref_vote.gtf")).read()      # committed output of the reference
    if os.path.exists(os.path.join(REF, "vote-transcripts")):
        assert ours == run([os.path.join(REF, "vote-transcripts"), "--lg", lst])
        assert run([os.path.join(BIN, "vote-transcripts"), "--lg", lst, "-d", "0.2"]) == \
            run([os.path.join(REF, "vote-transcripts"), "--lg", lst, "-d", "0.2"])


def _fused_vote(classes, name, tmp_path):
    d = stage_fixture(name, tmp_path)
    run([classes, "--lb", "bamlist", "-s", "subexon_combined.out", "-o", "new", "-p", "3", "--vote", "new_vote.gtf"] + FIXTURES[name], cwd=d)
    fused = open(os.path.join(d, "new_vote.gtf")).read()
    # the GTFs equal the reference's (other tests), so the reference's vote over the reference's GTFs is the answer
    if name == "example":
        assert fused == open(os.path.join(d, "ref_vote.gtf")).read()
    assert fused == run([os.path.join(BIN, "vote-transcripts"), "--lg", gtf_list(d, "ref")])
    assert fused == run([os.path.join(BIN, "vote-transcripts"), "--lg", gtf_list(d, "new")])


@pytest.mark.parametrize("name", ["dp_small", "example", "nh3_small"])
def test_fused_vote_on_cpu_shim(name, tmp_path):
    _fused_vote(os.path.join(SHIM, "classes_shim"), name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dp_small", "example", "nh3_small"])
def test_fused_vote_cuda(name, tmp_path):
    _fused_vote(os.path.join(BIN, "classes"), name, tmp_path)


def _driver_run(classes, tmp_path):
    if not os.path.exists(os.path.join(REF, "psiclass")) or shutil.which("perl") is None:
        pytest.skip("the reference driver (oracle/_ref/psiclass, perl) is not available")
    wd = os.path.join(str(tmp_path), "install")
    os.makedirs(wd)
    for f in os.listdir(REF):                                            # the reference install ...
        if f == "psiclass":                                              # abs_path($0) would follow a symlink
            shutil.copy(os.path.join(REF, f), os.path.join(wd, f))
        elif f not in ("classes", "vote-transcripts", "obj"):
            os.symlink(os.path.join(REF, f), os.path.join(wd, f))
    os.symlink(classes, os.path.join(wd, "classes"))                     # ... with our stage 3
    os.symlink(os.path.join(BIN, "vote-transcripts"), os.path.join(wd, "vote-transcripts"))  # ... and stage 4
    d = stage_fixture("example", tmp_path)
    out = subprocess.run(["perl", os.path.join(wd, "psiclass"), "--lb", "bamlist", "-o", "ex_", "-p", "2"], cwd=d,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert os.path.join(wd, "classes") in out.stdout                     # psiclass:410 really called the replacement
    assert gtf_body(os.path.join(d, "ex_vote.gtf")) == gtf_body(os.path.join(d, "ref_vote.gtf"))
    for s in range(2):
        assert gtf_body(os.path.join(d, "ex_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s))
    # stage 3 alone, resuming from the stage 0-2 files (psiclass:118-122, 395-420)
    out = subprocess.run(["perl", os.path.join(wd, "psiclass"), "--lb", "bamlist", "-o", "ex_", "--stage", "3"], cwd=d,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert gtf_body(os.path.join(d, "ex_vote.gtf")) == gtf_body(os.path.join(d, "ref_vote.gtf"))


def test_psiclass_driver_with_drop_in_on_cpu_shim(tmp_path):
    _driver_run(os.path.join(SHIM, "classes_shim"), tmp_path)


@pytest.mark.gpu
def test_psiclass_driver_with_drop_in_cuda(tmp_path):
    _driver_run(os.path.join(BIN, "classes"), tmp_path)


def test_perf_json_and_vote_depth_on_cpu_shim(tmp_path):
    """`PSC_PERF_JSON` (what bench.py reads) is valid JSON with the phase times and one entry per shard; `--vote-depth`
    is vote-transcripts' -d."""
    import json
    d = stage_fixture("dp_small", tmp_path)
    env = dict(os.environ, PSC_GPUS="2", PSC_PERF_JSON=os.path.join(d, "perf.json"))
    run([os.path.join(SHIM, "classes_shim"), "--lb", "bamlist", "-s", "subexon_combined.out", "-o", "new", "-p", "3",
         "--vote", "v.gtf", "--vote-depth", "0.2"], cwd=d, env=env)
    perf = json.load(open(os.path.join(d, "perf.json")))
    assert perf["gpus"] == 2 and len(perf["shards"]) == 2 and perf["samples"] == 3
    assert sum(s["components"] for s in perf["shards"]) == perf["components"]
    for k in ("decode_s", "graph_s", "device_s", "gtf_s", "vote_s", "total_s", "decode_MBps", "bam_bytes", "reads"):
        assert k in perf
    assert open(os.path.join(d, "v.gtf")).read() == run([os.path.join(BIN, "vote-transcripts"), "--lg", gtf_list(d, "new"), "-d", "0.2"])
