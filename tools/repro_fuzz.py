import os, random, sys
sys.path.insert(0,'/root/repo')
from psiclass_b200 import testkit
seed=int(sys.argv[1]); lib=sys.argv[2]
rng = random.Random(seed)
cfg = dict(seed=seed, n_genes=rng.choice([40, 80, 150]), n_samples=rng.choice([2, 3, 5, 8]),
           mean_pairs=rng.choice([5.0, 40.0, 150.0, 400.0]), exon_mu=rng.choice([1.6, 2.0, 2.4]),
           iso_lambda=rng.choice([1.5, 3.0, 5.0]), p_alt=rng.choice([0.1, 0.3, 0.45]), p_ir=rng.choice([0.0, 0.1]),
           p_noise=rng.choice([0.0, 0.1]), p_overlap=rng.choice([0.0, 0.15]), single_end=rng.choice([0, 0, 1]),
           stranded=rng.choice([0, 1]))
mode = rng.choice(["plain", "overlap", "chunks", "overlap+chunks"])
print(cfg, mode)
wl = testkit.Workload.synth(threads=8, **cfg)
rep = wl.check(lib, coop=int(sys.argv[3]) if len(sys.argv) > 3 else 32)
print(rep.passed, rep.summary()[-140:])
