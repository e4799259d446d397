import os, sys, tempfile
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from conftest import ROOT, gtf_body, run, stage_fixture
import pathlib
CLASSES = os.path.join(ROOT, "psiclass_b200", "bin", "classes")
for name in ["synth_small", "dp_small"]:
    tmp = pathlib.Path(tempfile.mkdtemp())
    d = stage_fixture(name, tmp)
    env = dict(os.environ, PSC_GPUS="2")
    out = run([CLASSES, "--lb", "bamlist", "-s", "subexon_combined.out", "-o", "new", "-p", "4"], cwd=d, env=env)
    n = len(open(os.path.join(d, "bamlist")).read().split())
    ok = all(gtf_body(os.path.join(d, "new_sample_%d.gtf" % s)) == gtf_body(os.path.join(d, "ref_sample_%d.gtf" % s)) for s in range(n))
    print(name, "PSC_GPUS=2 GTFs identical to reference:", ok)
