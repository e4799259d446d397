#!/usr/bin/env python
"""Parity of the CUDA engine against the oracle on a bench-like workload (default 6000 gene models x 10 samples;
the oracle is single-threaded, allow a few minutes).  usage: python tools/gpu_parity_large.py [n_genes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiclass_b200 import abi, testkit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
t0 = time.time()
wl = testkit.Workload.synth(seed=1001, n_chr=24, n_genes=n, n_samples=10, mean_pairs=60.0, threads=8)
rep = wl.check(abi.LIB_PATH, coop=32)
print("large parity:", "PASS" if rep.passed else "FAIL", rep.summary(), "%.0f s" % (time.time() - t0))
sys.exit(0 if rep.passed else 1)
