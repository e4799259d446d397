#!/usr/bin/env python
"""bench.py -- stage-3 throughput: subexon-graph components assembled per second (all samples).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine (libpsc.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU `classes`

Workload (BASELINE.json configs[1]): seeded synthetic 10-sample human-scale subexon graph
(~57k components from 60k gene models on 24 contigs, ~60 read pairs / component / sample), stage 3
only.  One "step" = one full pass of the hot path over the whole batch:

  value : psc_build_constraints + psc_solve with graph and alignment records resident in HBM
  e2e   : psc_set_reads (pinned host -> device) + build + solve + psc_get_transcripts (device -> host)

With N > 1 (torchrun, one rank per GPU) every rank processes its own workload of the same size
(weak scaling, no data-path collective); the time is the max over ranks.

File legs (rank 0): the SAME workload written as BAM files + combined subexon file and run through the
drop-in `psiclass_b200/bin/classes` (PSC_GPUS=N: ONE workload sharded by component over the N devices,
strong scaling) -- BAM on disk -> GTF on disk, wall clock around the process -> `e2e_files`, with a sha256 of
the GTF bodies.  `--impl reference` runs the unmodified reference `classes -p <cores>` on the same files
(same generator, same seed, same size) and prints the same digest, so the two arms are like for like.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# psc_solve drives several streams at once; with the default of 8 hardware queues some of them share one (set before
# CUDA is initialised; measured 641 -> 623 ms per step)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOAD = dict(seed=1001, n_chr=24, n_genes=60000, n_samples=10, mean_pairs=60.0)
CPU_SAMPLE = dict(seed=1001, n_chr=24, n_genes=1500, n_samples=10, mean_pairs=60.0)
# BASELINE.json configs[2..4] at a stated scale (file legs only: `python bench.py --config K --gpus N [--with-reference]`).
# Full size is 60000 gene models x 100 samples (configs[2]), 2000 dense components (configs[3]), 30 samples (configs[4]);
# the scales below keep the reference's CPU pass at minutes (its dense-gene pass runs ~20 s per component).
EXTRA_CONFIGS = {
    2: dict(name="configs[2] scaled: 100 samples, 1/10 of the graph (6000 gene models), ~40 read pairs/component/sample",
            synth=dict(seed=1002, n_chr=24, n_genes=6000, n_samples=100, mean_pairs=40.0), args=[]),
    3: dict(name="configs[3] scaled: 32 dense components of 200-600 subexons, 10 samples, depth x5, --maxDpConstraintSize -1",
            synth=dict(seed=1003, n_chr=4, n_genes=32, n_samples=10, dense=1, mean_pairs=300.0), args=["--maxDpConstraintSize", "-1"],
            ref_genes=4),
    4: dict(name="configs[4] scaled: 30 samples, 6000 gene models, rf-stranded flags, 15 % paralog pairs (NH=2 primary + secondary), --primaryParalog",
            synth=dict(seed=1004, n_chr=24, n_genes=6000, n_samples=30, mean_pairs=60.0, stranded=1, p_paralog=0.15),
            args=["--primaryParalog", "--stranded", "rf"]),
}
METRIC = "subexon-graph components assembled/sec (stage 3, all samples)"
UNIT = "components/s"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks line of the profiling recipe, sampled during the timed region."""

    def __init__(self, gpu):
        self.gpu = gpu
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # under load == the upper half of the samples
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def write_workload_files(cfg, cores, wl=None):
    """The workload as files: <dir>/subexon_combined.out, s<i>.bam, bamlist.  Returns (dir, #components, bytes)."""
    from psiclass_b200 import testkit
    own = wl is None
    if own:
        wl = testkit.Workload.synth(threads=cores, **cfg)
    d = tempfile.mkdtemp(prefix="psc_bench_")
    wl.write_files(d)
    ncomp = wl.n_comp
    if own:
        wl.close()
    nbytes = 0
    for f in os.listdir(d):  # warm the page cache
        with open(os.path.join(d, f), "rb") as fh:
            while True:
                b = fh.read(1 << 24)
                if not b:
                    break
                nbytes += len(b)
    return d, ncomp, nbytes


def gtf_digest(d, prefix, n_samples):
    """sha256 over the GTF bodies ('#' header lines carry argv and paths) of all samples, in order."""
    h = hashlib.sha256()
    lines = 0
    for s in range(n_samples):
        with open(os.path.join(d, "%s_sample_%d.gtf" % (prefix, s)), "rb") as f:
            for ln in f:
                if not ln.startswith(b"#"):
                    h.update(ln)
                    lines += 1
    return h.hexdigest(), lines


def run_classes(binary, d, prefix, threads, gpus=None, extra_env=None, extra_args=()):
    """One `classes` process on the files in d; returns (seconds, perf dict or None)."""
    env = dict(os.environ)
    perf_path = None
    if gpus is not None:
        env["PSC_GPUS"] = str(gpus)
        perf_path = os.path.join(d, prefix + "_perf.json")
        env["PSC_PERF_JSON"] = perf_path
    if extra_env:
        env.update(extra_env)
    cmd = [binary, "--lb", os.path.join(d, "bamlist"), "-s", os.path.join(d, "subexon_combined.out"),
           "-o", os.path.join(d, prefix), "-p", str(threads)] + list(extra_args)
    t0 = time.perf_counter()
    subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, env=env)
    dt = time.perf_counter() - t0
    perf = None
    if perf_path and os.path.exists(perf_path):
        try:
            perf = json.load(open(perf_path))
        except Exception:
            perf = None
    return dt, perf


def reference_binary():
    from psiclass_b200 import testkit
    ref = os.path.join(testkit.REF_DIR, "classes")
    if os.path.exists(ref):
        return ref, "reference"
    ref = os.path.join(testkit.REF_DIR, "psc_oracle")  # the CPU restatement (single-threaded)
    if os.path.exists(ref):
        return ref, "port"
    return None, None


def describe(cfg):
    return "%d gene models on %d contigs, seed %d, %d samples, ~%d read pairs/component/sample" % (
        cfg["n_genes"], cfg["n_chr"], cfg["seed"], cfg["n_samples"], int(cfg["mean_pairs"]))


def run_extra_config(a):
    """One workload as files, the drop-in `classes` on --gpus devices (sharded by component), optionally the reference."""
    ec = EXTRA_CONFIGS[a.config]
    cores = host_cores()
    syn = dict(ec["synth"])
    if a.genes:
        syn["n_genes"] = a.genes
    ours = os.path.join(ROOT, "psiclass_b200", "bin", "classes")
    d, ncomp, nbytes = write_workload_files(syn, cores)
    line = {"config": ec["name"], "workload": describe(syn) if "dense" not in syn else str(syn), "components": ncomp,
            "bam_bytes": nbytes, "n_gpus": a.gpus, "host_threads": cores, "unit": UNIT}
    try:
        t_new, perf = run_classes(ours, d, "new", cores, gpus=a.gpus, extra_args=ec["args"])
        digest, glines = gtf_digest(d, "new", syn["n_samples"])
        line["drop_in"] = {"seconds": t_new, "components_per_s": ncomp / t_new, "gtf_sha256": digest, "gtf_lines": glines}
        if perf:
            sh = perf.get("shards", [])
            line["drop_in"]["phases_s"] = {k: perf.get(k) for k in ("decode_s", "graph_s", "device_s", "gtf_s", "total_s")}
            if sh:
                pipe = max(x["upload_s"] + x["constraints_s"] + x["solve_s"] + x["download_s"] for x in sh)
                line["drop_in"]["device_pipeline"] = {"seconds": pipe, "components_per_s": ncomp / pipe,
                                                      "solve_s_per_shard": [x["solve_s"] for x in sh],
                                                      "components_per_shard": [x["components"] for x in sh]}
        ref, kind = reference_binary()
        if a.with_reference and ref is not None:
            rd, rn = d, ncomp
            if ec.get("ref_genes"):  # the reference needs ~20 s per dense component: a bounded sample of the same generator
                shutil.rmtree(d, ignore_errors=True)
                syn2 = dict(syn, n_genes=ec["ref_genes"])
                d, rn, _ = write_workload_files(syn2, cores)
                rd = d
                t2, _ = run_classes(ours, rd, "new", cores, gpus=1, extra_args=ec["args"])
                line["drop_in_on_reference_sample"] = {"components": rn, "seconds": t2}
            threads = cores if kind == "reference" else 1
            t_ref, _ = run_classes(ref, rd, "ref", threads, extra_args=ec["args"])
            rdig, _ = gtf_digest(rd, "ref", syn["n_samples"])
            ndig, _ = gtf_digest(rd, "new", syn["n_samples"])
            line["reference"] = {"kind": kind, "cores": threads, "components": rn, "seconds": t_ref, "components_per_s": rn / t_ref,
                                 "gtf_identical": rdig == ndig}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="psc")
    ap.add_argument("--genes", type=int, default=None, help="override the number of gene models (debugging)")
    ap.add_argument("--coop", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the BAM -> GTF legs through the `classes` binary")
    ap.add_argument("--ref-passes", type=int, default=2, help="timed passes of the reference arm (one pass is ~1-2 min)")
    ap.add_argument("--config", type=int, default=1, help="2, 3, 4: BAM -> GTF legs of BASELINE.json configs[2..4] at a stated scale")
    ap.add_argument("--with-reference", action="store_true", help="--config: also time the reference `classes` on the same files")
    a = ap.parse_args()
    if a.config in EXTRA_CONFIGS:
        return run_extra_config(a)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = host_cores()
    cfg = dict(WORKLOAD)
    cpu_cfg = dict(CPU_SAMPLE)
    if a.genes:
        cfg["n_genes"] = a.genes
        cpu_cfg["n_genes"] = min(a.genes, cpu_cfg["n_genes"])
    config = {"workload": "synthetic 10-sample human-scale subexon graph, stage 3 only (BASELINE.json configs[1]): " + describe(cfg),
              "l2": "inputs (alignment records + tables, several GB per GPU) are far larger than the 126 MB L2",
              "sharding": "components shard by rank with no collective; every rank assembles an identically generated "
                          "workload of this size (same seed), so the per-GPU work is exactly fixed as N grows"}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        ref, kind = reference_binary()
        if ref is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built"}))
            return 0
        # The SAME workload as the CUDA arm (same generator, seed and size), as files.  One pass of the reference
        # takes 1-2 minutes, so at most --ref-passes (2) timed passes run whatever --steps says, without a warm-up
        # pass (the page cache is warmed by reading the files once).
        passes = max(1, min(a.steps, a.ref_passes))
        threads = cores if kind == "reference" else 1
        d, ncomp, nbytes = write_workload_files(cfg, cores)
        try:
            times = []
            for _ in range(passes):
                dt, _ = run_classes(ref, d, "ref", threads)
                times.append(dt)
            digest, glines = gtf_digest(d, "ref", cfg["n_samples"])
        finally:
            shutil.rmtree(d, ignore_errors=True)
        mean = sum(times) / len(times)
        value = ncomp / mean
        sample = ("the whole workload (%s): %d components, %.2f GB of BAM + subexon file on local disk, page cache warm, "
                  "`classes -p %d` (BAM decode + stage 3 + GTF), %d timed passes of %.1f s, no warm-up pass"
                  % (describe(cfg), ncomp, nbytes / 1e9, threads, passes, mean))
        config["reference_passes"] = passes
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
                "steps": passes, "warmup": 0, "ms_per_step": 1000.0 * mean,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "e2e_files": {"value": value, "unit": UNIT, "seconds": mean, "components": ncomp, "bam_bytes": nbytes,
                              "gtf_sha256": digest, "gtf_lines": glines},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ CUDA arm
    import torch
    from psiclass_b200 import abi, testkit

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side wait while rank 0 runs the file legs

    # Same seed on every rank: the step time of this path is set by the single heaviest component of the
    # workload, which varies by +-50 % between seeds; identical replicas keep the per-GPU work fixed,
    # which is what weak scaling is meant to measure.
    cfg_rank = dict(cfg)
    wl = testkit.Workload.synth(threads=max(1, cores // max(1, world)), **cfg_rank)
    S, ncomp = wl.n_samples, wl.n_comp
    eng = abi.Engine(device=local_rank)
    prm = wl.params()
    eng.call("psc_set_params", C.byref(prm))
    eng.call("psc_set_coop_width", a.coop)
    g = wl.graph()
    eng.call("psc_set_graph", C.byref(g))
    reads = [wl.reads(s) for s in range(S)]
    # pin the host copies of the alignment records (the H2D source of the e2e path)
    cudart = torch.cuda.cudart()
    h2d = 0
    for r in reads:
        n, ns = r.n_reads, r.n_segs
        for ptr, nbytes in ((r.pos, 4 * n), (r.mate_pos, 4 * n), (r.flags, n), (r.nh, n), (r.name_key, 8 * n),
                            (r.seg_off, 4 * (n + 1)), (r.seg_a, 4 * ns), (r.seg_b, 4 * ns)):
            if nbytes:
                cudart.cudaHostRegister(C.cast(ptr, C.c_void_p).value, nbytes, 0)
            h2d += nbytes

    def upload():
        for s in range(S):
            eng.call("psc_set_reads", s, C.byref(reads[s]))

    out_bufs = {}

    def download():
        nt, ne = C.c_int64(), C.c_int64()
        eng.call("psc_transcript_counts", C.byref(nt), C.byref(ne))
        nt, ne = nt.value, ne.value
        if out_bufs.get("cap", (0, 0)) < (nt, ne):
            import numpy as np
            out_bufs["cap"] = (nt, ne)
            out_bufs["arrs"] = [np.empty(nt + 1, np.int32) for _ in range(4)] + [np.empty(nt + 1, np.int8)] + \
                [np.empty(nt + 2, np.int64)] + [np.empty(ne + 1, np.int32) for _ in range(2)] + \
                [np.empty(nt + 1, np.float64) for _ in range(2)]
        arrs = out_bufs["arrs"]
        t = abi.psc_transcripts()
        t.comp, t.sample, t.chr, t.gene_id = [x.ctypes.data_as(abi.i32p) for x in arrs[0:4]]
        t.strand = arrs[4].ctypes.data_as(abi.i8p)
        t.exon_off = arrs[5].ctypes.data_as(abi.i64p)
        t.exon_start, t.exon_end = [x.ctypes.data_as(abi.i32p) for x in arrs[6:8]]
        t.fpkm, t.cov = [x.ctypes.data_as(abi.f64p) for x in arrs[8:10]]
        eng.call("psc_get_transcripts", C.byref(t))
        return nt, nt * 34 + ne * 8

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        eng.call("psc_build_constraints")
        eng.call("psc_solve")

    def step_e2e():
        upload()
        eng.call("psc_build_constraints")
        eng.call("psc_solve")
        return download()

    upload()
    for _ in range(a.warmup):
        step_resident()
    # ---- timed: inputs resident ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, kms, kcnt, kbytes, launches = 0.0, [0.0] * 12, [0] * 12, [0] * 12, 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_resident()
        tm = eng.timing()
        dev_ms += tm.constraints_ms + tm.solve_ms
        launches += tm.kernel_launches
        for k in range(12):
            kms[k] += tm.kernel_ms[k]
            kcnt[k] += tm.kernel_count[k]
            kbytes[k] = tm.kernel_bytes[k]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # the step includes small host-side sizing passes between launches, so the step time is the
    # synchronised wall time around the K steps (device events cover the kernels only)
    t_res = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    wall_max = float(t_res.item())
    # ---- timed: end to end through the C ABI with host buffers ----
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    ntx = 0
    for _ in range(a.steps):
        ntx, d2h = step_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t0
    t_e = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_max = float(t_e.item())
    tot_comp = torch.tensor([float(ncomp)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot_comp, op=dist.ReduceOp.SUM)
    total_components = float(tot_comp.item())

    if rank == 0:
        names = ["group_ranges", "build_constraints", "prologue", "enumerate", "dp_units", "units_pick1", "emit", "dp_merge",
                 "units_pick2_refine_em", "units_rescue", "dp_program", "dp_program_thread_per_unit"]
        dom = max(range(12), key=lambda k: kms[k])
        peak, peak_src = measured_peak()
        per_launch_ms = kms[dom] / max(1, kcnt[dom])
        # algorithmic bytes are modelled per step for kernels 1 and 5; per launch = per step / launches per step
        launches_per_step = max(1, kcnt[dom] // max(1, a.steps))
        alg = kbytes[dom] / launches_per_step
        achieved = (alg / 1e9) / (per_launch_ms / 1e3) if per_launch_ms > 0 and alg > 0 else 0.0
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: read from the JSON that
        # tools/gpu/ncu_traffic.py wrote from an `ncu --set full` capture of this workload (profiles/), else null
        def traffic_of(kernel):
            try:
                ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
                return float(ent["dram_bytes_per_launch"]) if ent and not a.genes else None
            except Exception:
                return None
        traffic = traffic_of("k_" + names[dom])
        line = {
            "metric": METRIC, "value": total_components * a.steps / wall_max, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * wall_max / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": total_components * a.steps / e2e_max, "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_" + names[dom], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "ms_per_launch": per_launch_ms},
            # the one streaming kernel of the path (alignment records in, constraint tables out), reported beside the
            # dominant (latency-bound) kernel: algorithmic bytes of the step / its launch time / the measured HBM peak
            "roofline_streaming": (lambda ms, by: {"bound": "hbm", "kernel": "k_build_constraints", "achieved": (by / 1e9) / (ms / 1e3) if ms > 0 else 0.0,
                                                     "peak": peak, "unit": "GB/s", "frac": ((by / 1e9) / (ms / 1e3) / peak) if ms > 0 and peak else None,
                                                     "traffic": traffic_of("k_build_constraints"), "algorithmic_bytes_per_launch": by,
                                                     "ms_per_launch": ms})(kms[1] / max(1, kcnt[1]), float(kbytes[1])),
            "kernel_ms_per_step": {names[k]: kms[k] / a.steps for k in range(12) if kcnt[k]},
            "device_ms_per_step": dev_ms / a.steps,
            "components_per_rank": ncomp, "transcripts_per_rank": int(ntx),
            "reads_per_rank": int(sum(r.n_reads for r in reads)),
        }
        eng.close()
        ours = os.path.join(ROOT, "psiclass_b200", "bin", "classes")
        ref, kind = reference_binary()
        if not a.no_cpu_baseline and world == 1 and ref is not None:
            # the reference's CPU path on the box's host cores, on a bounded sample of the same workload (same generator
            # and seed, 1/40 of the gene models); the drop-in runs on the same files and the GTFs are compared
            try:
                d, nc_s, nb_s = write_workload_files(cpu_cfg, cores)
                try:
                    threads = cores if kind == "reference" else 1
                    t_ref, _ = run_classes(ref, d, "ref", threads)
                    t_new, _ = run_classes(ours, d, "new", cores, gpus=1)
                    same = gtf_digest(d, "ref", cpu_cfg["n_samples"]) == gtf_digest(d, "new", cpu_cfg["n_samples"])
                finally:
                    shutil.rmtree(d, ignore_errors=True)
                line["cpu_baseline"] = {
                    "value": nc_s / t_ref, "unit": UNIT, "cores": threads, "kind": kind,
                    "sample": "%d components (%s): BAM files on local disk, page cache warm, `classes -p %d` (BAM decode + "
                              "stage 3 + GTF), one pass of %.2f s; the whole workload is timed by `--impl reference`"
                              % (nc_s, describe(cpu_cfg), threads, t_ref),
                    "drop_in_same_files": {"seconds": t_new, "components_per_s": nc_s / t_new, "gtf_identical": bool(same)}}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": kind, "sample": "failed: %s" % e}
        if not a.no_files:
            # BAM on disk -> GTF on disk through the drop-in binary: ONE workload of the full size, sharded by
            # component over the N devices (PSC_GPUS=N), no collective, final gather + global sort on the host
            try:
                d, nc_f, nb_f = write_workload_files(cfg, cores, wl)
                try:
                    t_new, perf = run_classes(ours, d, "new", cores, gpus=world)
                    digest, glines = gtf_digest(d, "new", cfg["n_samples"])
                finally:
                    shutil.rmtree(d, ignore_errors=True)
                ef = {"value": nc_f / t_new, "unit": UNIT, "seconds": t_new, "components": nc_f, "bam_bytes": nb_f,
                      "n_gpus": world, "scaling": "strong", "host_threads": cores, "gtf_sha256": digest, "gtf_lines": glines}
                if perf:
                    sh = perf.get("shards", [])
                    ef["phases_s"] = {k: perf.get(k) for k in ("decode_s", "graph_s", "device_s", "gtf_s", "total_s")}
                    ef["decode_MBps"] = perf.get("decode_MBps")
                    if sh:
                        pipe = max(x["upload_s"] + x["constraints_s"] + x["solve_s"] + x["download_s"] for x in sh)
                        ef["device_pipeline"] = {
                            "seconds": pipe, "components_per_s": nc_f / pipe if pipe > 0 else None,
                            "what": "max over the shards of upload + constraints + solve + download (one workload, sharded)",
                            "shards": sh}
                line["e2e_files"] = ef
            except Exception as e:
                line["e2e_files"] = {"value": None, "unit": UNIT, "error": str(e)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
