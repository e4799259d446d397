"""world_size-2 `gloo` test of the multi-rank plumbing bench.py uses: every rank assembles its own
shard (by component, no data-path collective); only the component totals and the max step time are
reduced.  Runs on the CPU (shim engine) -- the N>1 CUDA path is the same code with backend nccl."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psiclass_b200 import testkit
    wl = testkit.Workload.synth(seed=51 + 7919 * rank, n_genes=20, n_samples=2, threads=1)
    rep = wl.check(testkit.SHIM_PATH, coop=1)
    ok = torch.tensor([1.0 if rep.passed else 0.0], dtype=torch.float64)
    comps = torch.tensor([float(wl.n_comp)], dtype=torch.float64)
    t = torch.tensor([0.5 + rank], dtype=torch.float64)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.all_reduce(comps, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((ok.item(), comps.item(), t.item(), wl.n_comp))
    dist.destroy_process_group()


def test_two_ranks_shard_by_component():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, comps, t, n0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok == 1.0
    assert comps > n0          # the other rank's (different) workload was counted
    assert t == 1.5            # max over ranks
