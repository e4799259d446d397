#!/usr/bin/env python
"""Randomised parity sweep of the CUDA engine against the oracle (test infrastructure; run on a GPU box).
usage: python tools/gpu_fuzz.py [first_seed] [count] [shim]
With a third argument `shim` the CPU harness of the kernels (tests/shim) is swept instead (no GPU needed)."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiclass_b200 import abi, testkit

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
shim = len(sys.argv) > 3 and sys.argv[3] == "shim"
bad = 0
t0 = time.time()
for seed in range(first, first + count):
    rng = random.Random(seed)
    cfg = dict(seed=seed, n_genes=rng.choice([40, 80, 150]), n_samples=rng.choice([2, 3, 5, 8]),
               mean_pairs=rng.choice([5.0, 40.0, 150.0, 400.0]), exon_mu=rng.choice([1.6, 2.0, 2.4]),
               iso_lambda=rng.choice([1.5, 3.0, 5.0]), p_alt=rng.choice([0.1, 0.3, 0.45]), p_ir=rng.choice([0.0, 0.1]),
               p_noise=rng.choice([0.0, 0.1]), p_overlap=rng.choice([0.0, 0.15]), single_end=rng.choice([0, 0, 1]),
               stranded=rng.choice([0, 1]), p_nh3=rng.choice([0.0, 0.0, 0.1, 0.3]), p_paralog=rng.choice([0.0, 0.0, 0.15]))
    max_dp = rng.choice([7, 7, 7, 3, -1]) if cfg["exon_mu"] < 2.4 else 7
    prim = rng.choice([0, 0, 1])
    env = {}
    mode = rng.choice(["plain", "overlap", "chunks", "overlap+chunks"])
    for k in ("PSC_FORCE_OVERLAP", "PSC_MEM_BUDGET_MB"):
        os.environ.pop(k, None)
    if "overlap" in mode:
        os.environ["PSC_FORCE_OVERLAP"] = "1"
    if "chunks" in mode:
        os.environ["PSC_MEM_BUDGET_MB"] = "4"
    wl = testkit.Workload.synth(threads=8, **cfg)
    rep = wl.check(testkit.SHIM_PATH, coop=1, max_dp=max_dp, primary_paralog=prim) if shim else wl.check(abi.LIB_PATH, coop=32, max_dp=max_dp, primary_paralog=prim)
    ok = bool(rep.passed)
    bad += 0 if ok else 1
    print("%s seed %d %-14s maxDp %d prim %d %s | %s" % ("ok  " if ok else "FAIL", seed, mode, max_dp, prim, rep.summary()[-110:], cfg), flush=True)
print("fuzz: %d configurations, %d failed, %.0f s" % (count, bad, time.time() - t0))
sys.exit(1 if bad else 0)
