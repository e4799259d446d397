import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Everything native is built once per session (no-op when up to date)."""
    import __graft_entry__ as ge
    need = [os.path.join(ROOT, "build", "libpsc_testkit.so"), os.path.join(ROOT, "psiclass_b200", "libpsc.so"),
            os.path.join(ROOT, "oracle", "_ref", "psc_oracle"), os.path.join(ROOT, "tests", "shim", "_build", "libpsc_shim.so")]
    if not all(os.path.exists(p) for p in need):
        ge.build()
    return True


def stage_fixture(name, tmp_path):
    """Copies a golden fixture into tmp_path with an absolute bamlist; returns the directory."""
    src = os.path.join(GOLDEN, name)
    dst = os.path.join(str(tmp_path), name)
    shutil.copytree(src, dst)
    with open(os.path.join(src, "bamlist")) as f:
        bams = [ln.strip() for ln in f if ln.strip()]
    with open(os.path.join(dst, "bamlist"), "w") as f:
        for b in bams:
            f.write(os.path.join(dst, os.path.basename(b)) + "\n")
    return dst


def gtf_body(path):
    with open(path) as f:
        return [ln for ln in f if not ln.startswith("#")]


def run(cmd, cwd=None, env=None, timeout=900):
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0, "command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:])
    return r.stdout
