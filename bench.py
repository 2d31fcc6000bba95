#!/usr/bin/env python3
"""bench.py — k-mer counting throughput of the environment-finder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

Metric (BASELINE.json / SURVEY.md 8d): Gbases/s = input bases / time of [parse + extract + count + coverage
compaction].  One "step" = one pass of that path over the synthetic metagenome of the named workload:
  value  — reads already resident in HBM when the timed region starts (device generator, include/mc_synth.h)
  e2e    — the same job through the C ABI with HOST buffers (mc_count_bytes): H2D copies and the D2H read of
           the result (size of the coverage-filtered table) inside the timed region
Inputs are >> L2 (15.75 GB of FASTQ per step at cfg2), so no explicit L2 flush is needed between steps.
The roofline entry is for the counting pipeline (k_extract + k_bin_scatter + k_bin_scatter2 + k_bin_agg: the kernels
that together do what SURVEY.md 8d's per-read byte figure describes);
cpu_baseline times the oracle's multi-threaded C++ restatement of the reference on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# SURVEY.md 8d workloads.  cfg2 = BASELINE.json configs[1], the configuration the metric is quoted on.
WORKLOADS = {
    "cfg1": dict(reads=1_000_000, genome_unit=100_000, n_targets=1, k=31, coverage=5, err_ppm=5000, repeat_len=0,
                 desc="environment-finder k=31 coverage=5, 1M synthetic 150bp FASTQ reads, 10-genome mock community (5.5 Mbp), 1 target"),
    "cfg2": dict(reads=50_000_000, genome_unit=1_000_000, n_targets=20, k=31, coverage=5, err_ppm=5000, repeat_len=0,
                 desc="environment-finder k=31 coverage=5, 50M synthetic 150bp FASTQ reads, 10-genome community (55 Mbp), 20 targets"),
}
READ_LEN = 150
FMT_FASTQ = 1
# algorithmic bytes per read = record bytes + (L-k+1) * 64 (one 32-B sector read + written back per k-mer instance)
def algorithmic_bytes_per_read(k, record_bytes):
    return record_bytes + (READ_LEN - k + 1) * 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_for(oracle, w):
    return oracle.synth(n_genomes=10, genome_unit=w["genome_unit"], read_len=READ_LEN, n_targets=w["n_targets"], target_len=900,
                        repeat_len=w["repeat_len"], err_ppm=w["err_ppm"], n_ppm=100, bangq_ppm=100, fasta=False)


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Java reference
    cannot run here (no JVM, compiled ITMO jar missing) so this is the oracle's C++ restatement with the
    reference's structure (serialized parser + P workers over lock-striped maps), kind = "port"."""
    if rank != 0:
        return
    import oracle_ffi
    o = oracle_ffi.load()
    cores = os.cpu_count() or 1
    s = synth_for(o, w)
    sample_reads = min(w["reads"], args.ref_sample_reads)
    data = s.generate(0, sample_reads, threads=cores)
    times = []
    bases = 0
    for i in range(args.warmup + args.steps):
        secs, bases, distinct = o.bench_count(data, FMT_FASTQ, w["k"], cores)
        if i >= args.warmup:
            times.append(secs)
    total = sum(times)
    val = bases * len(times) / total / 1e9
    sample = f"first {sample_reads} reads of the workload ({data.size / 1e6:.0f} MB FASTQ) per step, parse+count only, multi-threaded C++ restatement"
    line = {"metric": "kmer_count_throughput", "value": round(val, 5), "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * total / len(times), 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": w["name"], "description": w["desc"], "k": w["k"], "coverage": w["coverage"], "reads_per_step": sample_reads},
            "cpu_baseline": {"value": round(val, 5), "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 5), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="debug only: override reads per step (result is then NOT the named workload)")
    ap.add_argument("--chunk-mib", type=int, default=1024)
    ap.add_argument("--ref-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-env", action="store_true")
    ap.add_argument("--maxkmers", type=int, default=100000)
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=VALUE", help="mc_ctx_set_tuning knob (part_bits, batch_keys, segment_keys)")
    ap.add_argument("--force-sharded", action="store_true", help="profiling: run the N > 1 code path on one rank")
    ap.add_argument("--exchange", default="merge", choices=["merge", "peer", "nccl"], help="N > 1: how the ranks' k-mers reach their owners (metacherchant_b200/dist.py)")
    ap.add_argument("--hint-div", type=int, default=0, help="table capacity hint = bases / this; 0 = none: the library sizes the table from a 1/64 sample of the first batch")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload], name=args.workload)
    if args.reads:
        w["reads"] = args.reads
        w["name"] = f"{args.workload}-debug-{args.reads}-reads"

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_ffi   # only for the synthetic-workload parameters and the cpu_baseline leg
    from metacherchant_b200 import Context

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures the CUDA path only (no CPU fallback)"}), flush=True)
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    o = oracle_ffi.load()
    s = synth_for(o, w)
    params = s.params_bytes()
    rb = s.record_bytes
    reads = w["reads"]          # per GPU (weak scaling: every rank counts its own slice of the read stream)
    nbytes = reads * rb
    k, cov = w["k"], w["coverage"]
    bases_per_step = reads * READ_LEN

    ctx = Context(local_rank, chunk_bytes=args.chunk_mib << 20)
    for kv in args.tune:
        name, v = kv.split("=")
        ctx.set_tuning(name, int(v))
    d_reads = ctx.device_alloc(nbytes)
    first = rank * reads
    GEN = 4_000_000
    for off in range(0, reads, GEN):   # device generator, not timed
        n = min(GEN, reads - off)
        ctx.synth_generate_device(params, first + off, n, d_reads + off * rb)
    # no rehash when memory allows: distinct k-mers <= k-mer instances <= bases (what mc_count_files does from the file size)
    hint = bases_per_step // args.hint_div if args.hint_div > 0 else 0   # expected distinct k-mers; the table still grows (rehash) on demand beyond this

    sharded = world > 1 or args.force_sharded
    if sharded:
        from metacherchant_b200.dist import ShardedCounter
        counter = ShardedCounter(ctx, k, world, rank, capacity_hint=hint, exchange=args.exchange)

    def step_device():
        if sharded:
            return counter.count_device_step(d_reads, nbytes, FMT_FASTQ, cov, batch_bytes=args.chunk_mib << 20, record_bytes=rb)
        t = ctx.new_table(k, capacity_hint=hint)
        t.count_device(d_reads, nbytes, FMT_FASTQ)
        solid = t.filter(cov)
        n_solid = solid.size()
        stage = ctx.stage_ms()
        solid.free()
        t.free()
        return n_solid, stage

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    ctx.event_record(0)
    t0 = time.perf_counter()
    kernel_ms = 0.0
    n_solid = 0
    stages = {}
    for _ in range(args.steps):
        n_solid, stage = step_device()
        # the counting pipeline = extract (k-mers -> hashed keys) + scatter (keys -> bins) + count (bins -> table):
        # together they do what SURVEY 8d's per-read byte figure describes (record read + one table update per k-mer)
        kernel_ms += stage["extract"] + stage["scatter"] + stage["count"]
        for name, v in stage.items():
            stages[name] = stages.get(name, 0.0) + v
    ctx.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    # max over ranks
    t_ms = max(dev_ms, 0.0)
    if world > 1:
        tt = torch.tensor([t_ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, kernel_ms = tt.tolist()
    value = bases_per_step * world * args.steps / (t_ms / 1e3) / 1e9

    # ---- reads -> environments (the other half of BASELINE.json's metric), once, outside the timed region -----------
    env = None
    if world == 1 and not args.no_env:
        import shutil
        t = ctx.new_table(k, capacity_hint=hint)
        t.count_device(d_reads, nbytes, FMT_FASTQ)
        solid = t.filter(cov)
        out_dir = tempfile.mkdtemp(prefix="mc_env_")
        try:
            fa = os.path.join(out_dir, "targets.fasta")
            with open(fa, "w") as f:
                f.write(s.targets_fasta())
            targets = ctx.load_targets(fa)
            ctx.sync()
            t_env = time.perf_counter()
            found = ctx.find_environments(solid, [x[1] for x in targets], [x[0] for x in targets], min_cov=cov, max_kmers=args.maxkmers,
                                          both_dirs=False, out_dir=os.path.join(out_dir, "out"))
            env_s = time.perf_counter() - t_env
            n_env_kmers = 0
            for name, _ in targets:
                g = os.path.join(out_dir, "out", name, "env.txt")
                if os.path.exists(g):
                    with open(g) as f:
                        n_env_kmers += sum(1 for _ in f)
            env = {"targets": len(targets), "found": int(sum(found)), "maxkmers": args.maxkmers, "bothdirs": False, "env_kmers": n_env_kmers,
                   "env_ms": round(1000 * env_s, 1), "what": "BFS + unitigs + graph.gfa/env.txt/seqs.fasta/tsvs for all targets (mc_env_find, files written to a tmpfs dir)",
                   "reads_to_env_s": round(t_ms / args.steps / 1e3 + env_s, 3)}
        finally:
            shutil.rmtree(out_dir, ignore_errors=True)
            solid.free()
            t.free()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------
    e2e = None
    if not args.no_e2e and world == 1:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        ctx.d2h(host.numpy(), d_reads)
        ctx.device_free(d_reads)
        d_reads = None

        def step_host():
            t = ctx.new_table(k, capacity_hint=hint)
            t.count_host_ptr(host.data_ptr(), nbytes, FMT_FASTQ)
            solid = t.filter(cov)
            n = solid.size()       # the D2H read of the step's result
            solid.free()
            t.free()
            return n
        for _ in range(min(args.warmup, 3)):
            step_host()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            n_e2e = step_host()
        barrier()
        e2e_s = time.perf_counter() - t1
        assert n_e2e == n_solid, f"e2e result differs from the device-resident result: {n_e2e} vs {n_solid}"
        e2e = {"value": round(bases_per_step * args.steps / e2e_s / 1e9, 4), "unit": "Gbases/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": 8, "ms_per_step": round(1000 * e2e_s / args.steps, 2)}

    # ---- roofline of the dominant kernel (k_extract) -----------------------------------------------
    peak, peak_src = load_peaks()
    alg_bytes = algorithmic_bytes_per_read(k, rb) * reads * args.steps   # per rank
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None
    roofline = {"kernel": "counting pipeline: k_extract + k_bin_scatter + k_bin_scatter2 + k_bin_agg", "bound": "hbm", "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_read": algorithmic_bytes_per_read(k, rb), "kernel_ms_per_step": round(kernel_ms / args.steps, 2),
                "stage_ms_per_step": {name: round(v / args.steps, 2) for name, v in stages.items()}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            per_read = tj.get("pipeline_dram_bytes_per_read")
            # measured DRAM bytes of the four pipeline kernels per read (ncu over one full cfg2 step, profiles/r1d_NOTES.md) x reads of one step
            roofline["traffic"] = int(per_read * reads) if per_read else None
            roofline["traffic_note"] = ("ncu dram__bytes_read.sum + dram__bytes_write.sum of every launch of k_extract + k_bin_scatter + k_bin_scatter2 + k_bin_agg "
                                        "in one step (profiles/r1d_launches.csv), per read x reads per step")
            # what the pipeline really moved per second of its own time: the HBM utilisation behind `frac`
            if per_read and kernel_ms > 0:
                roofline["dram_gbs_measured_traffic"] = round(per_read * reads * args.steps / (kernel_ms / 1e3) / 1e9, 1)
        except Exception:
            pass
    # per-stage share of the pipeline (CUDA events on the library's stream): scatter = both bin levels, count = aggregation + table updates
    if kernel_ms > 0:
        roofline["pipeline_share"] = {name: round(stages.get(name, 0.0) / kernel_ms, 3) for name in ("extract", "scatter", "count")}

    # ---- CPU baseline (rank 0, N = 1): the oracle's multi-threaded restatement on a bounded sample --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_reads = min(reads, args.cpu_sample_reads)
        data = s.generate(0, sample_reads, threads=cores)
        secs, b, _ = o.bench_count(data, FMT_FASTQ, k, cores)
        cpu = {"value": round(b / secs / 1e9, 5), "unit": "Gbases/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_reads} reads of the workload ({data.size / 1e6:.0f} MB FASTQ), parse+count, C++ restatement of the Java reference, {secs:.1f} s"}

    if rank == 0:
        line = {"metric": "kmer_count_throughput", "value": round(value, 4), "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(t_ms / args.steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int64", "data": "synthetic",
                "config": {"workload": w["name"], "description": w["desc"], "k": k, "coverage": cov, "reads_per_gpu_per_step": reads,
                           "fastq_bytes_per_gpu_per_step": nbytes, "l2": "inputs (>= 0.3 GB per step) larger than the 126 MB L2; no explicit flush",
                           "chunk_mib": args.chunk_mib, "table": (f"pre-sized for bases/{args.hint_div} distinct k-mers, grows by rehash beyond" if args.hint_div > 0 else
                                     "no capacity hint: sized by the library from a counted 1/64 sample of the hash space (load 0.5), grows by rehash beyond"),
                           "solid_kmers": int(n_solid), "wall_s": round(wall, 3)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "env": env}
        print(json.dumps(line), flush=True)
    if sharded:
        counter.close_exchange()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
