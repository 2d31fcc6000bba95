#!/usr/bin/env python3
"""bench.py — k-mer counting throughput and reads -> environment time of the environment-finder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

Metric (BASELINE.json / SURVEY.md 8d): Gbases/s = input bases / time of [parse + extract + count + coverage
compaction].  One "step" = one pass of that path over the synthetic metagenome of the named workload:
  value  — reads already resident in HBM when the timed region starts (device generator, include/mc_synth.h)
  e2e    — the same job through the C ABI with HOST buffers (mc_count_bytes): H2D copies and the D2H read of
           the result (size of the coverage-filtered table) inside the timed region
Inputs are >> L2 (15.75 GB of FASTQ per step at cfg2), so no explicit L2 flush is needed between steps.
roofline.frac follows SURVEY.md 8d literally: algorithmic bytes per read x reads / the WHOLE step time (parse, extract,
both bin levels, aggregation, table growth, coverage compaction and every host gap between them); `frac_pipeline` is
the same bytes over the device time of the extract + scatter + count kernels alone.
cpu_baseline times the oracle's multi-threaded C++ restatement of the reference on a bounded sample (counting) and on
the workload's targets (BFS + unitigs + writers) with all host cores.

Workloads: cfg1 / cfg2 = BASELINE.json configs[0] / [1] (cfg2 is the one the metric is quoted on and
the default); cfg4 = configs[3] (k = 23, 200 M reads, 500 targets, --bothdirs, --maxkmers hit); cfg5 = configs[4] (1 B
reads in total at a 0.1 % error rate, strong scaling over the ranks).  cfg4 / cfg5 do not fit HBM as FASTQ text: their
reads are generated on the device batch by batch and only the counting calls are timed (CUDA events), see `config.timing`.
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
    "cfg1": dict(reads=1_000_000, genome_unit=100_000, n_targets=1, k=31, coverage=5, err_ppm=5000, repeat_len=0, bothdirs=False, streamed=False, scaling="weak",
                 desc="environment-finder k=31 coverage=5, 1M synthetic 150bp FASTQ reads, 10-genome mock community (5.5 Mbp), 1 target"),
    "cfg2": dict(reads=50_000_000, genome_unit=1_000_000, n_targets=20, k=31, coverage=5, err_ppm=5000, repeat_len=0, bothdirs=False, streamed=False, scaling="weak",
                 desc="environment-finder k=31 coverage=5, 50M synthetic 150bp FASTQ reads, 10-genome community (55 Mbp), 20 targets"),
    "cfg4": dict(reads=200_000_000, genome_unit=1_000_000, n_targets=500, k=23, coverage=5, err_ppm=5000, repeat_len=3000, bothdirs=True, streamed=True, scaling="strong",
                 desc="environment-finder k=23 coverage=5 --bothdirs True --maxkmers 100000, 200M synthetic 150bp FASTQ reads, 10-genome community (55 Mbp), "
                      "500 targets each next to a planted 3 kb shared repeat (deep, wide frontiers)"),
    "cfg5": dict(reads=1_000_000_000, genome_unit=1_000_000, n_targets=20, k=31, coverage=5, err_ppm=1000, repeat_len=0, bothdirs=False, streamed=True, scaling="strong",
                 desc="k-mer counting + coverage filter, 1B synthetic 150bp FASTQ reads IN TOTAL (split over the ranks), 0.1 % error rate, 10-genome community (55 Mbp)"),
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
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe), read through NVML in a thread of this
    process.  The recipe's `nvidia-smi -lms 200` child process was measured to cost 10-30 ms per 165 ms step on these boxes
    (profiles/r2_NOTES.md: every poll re-enumerates the devices and holds driver locks the launching thread needs), which made
    the number a function of the sampler; two NVML calls per poll do not.  Falls back to nvidia-smi when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, interval_s=0.25):
        self.gpu = gpu_index
        self.interval = interval_s
        self.p = None
        self.f = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.sm, self.mx, self.reasons = [], [], set()
        self.how = None

    def _poll_nvml(self, nv, h):
        names = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def start(self):
        try:
            if self.interval is None:
                raise RuntimeError("nvidia-smi sampler requested")
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES is not set by the launchers used here: NVML index == CUDA index
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.how = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.how = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "1000", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            self.how = "nvidia-smi -lms 1000"
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=5)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                    "samples": len(sm), "how": self.how}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sampler available"]}
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
                "samples": len(sm), "how": self.how}


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
    sample = (f"first {sample_reads} reads of the workload ({data.size / 1e6:.0f} MB FASTQ) per step (a SAMPLE of the GPU arm's {w['reads']} reads: throughput "
              f"is per base), parse+count only, multi-threaded C++ restatement; best step {min(times):.2f} s, median {sorted(times)[len(times) // 2]:.2f} s")
    line = {"metric": "kmer_count_throughput", "value": round(val, 5), "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * total / len(times), 2), "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "int64",
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
    ap.add_argument("--targets", type=int, default=0, help="debug only: override the number of targets")
    ap.add_argument("--chunk-mib", type=int, default=1024)
    ap.add_argument("--batch-reads", type=int, default=25_000_000, help="streamed workloads (cfg4, cfg5): reads generated and counted per batch")
    ap.add_argument("--ref-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-env", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="experiment: no clock sampler beside the timed region")
    ap.add_argument("--clocks", default="nvml", choices=["nvml", "smi"], help="clock sampler: NVML thread in this process (default) or an nvidia-smi child process")
    ap.add_argument("--maxkmers", type=int, default=100000)
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=VALUE", help="mc_ctx_set_tuning knob (part_bits, batch_keys, segment_keys)")
    ap.add_argument("--force-sharded", action="store_true", help="profiling: run the N > 1 code path on one rank")
    ap.add_argument("--exchange", default="merge", choices=["merge", "peer", "nccl"], help="N > 1: how the ranks' k-mers reach their owners (metacherchant_b200/dist.py)")
    ap.add_argument("--hint-div", type=int, default=0, help="table capacity hint = bases / this; 0 = none: the library sizes the table from a 1/64 sample of the first batch")
    ap.add_argument("--cpu-env-targets", type=int, default=20, help="CPU baseline of the environment stage: at most this many targets (the first ones)")
    ap.add_argument("--keep-env-dir", default="", help="write the environments here instead of a temporary directory (kept)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload], name=args.workload)
    if args.reads:
        w["reads"] = args.reads
        w["name"] = f"{args.workload}-debug-{args.reads}-reads"
    if args.targets:
        w["n_targets"] = args.targets
        w["name"] += f"-{args.targets}-targets"

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_ffi   # only for the synthetic-workload parameters and the cpu_baseline leg
    from metacherchant_b200 import Context

    try:
        os.nice(-10)          # a shared box: keep the launching thread ahead of other tenants' work (ignored without the privilege)
    except Exception:
        pass
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
    strong = w["scaling"] == "strong"
    # weak scaling: every rank counts its own w["reads"] reads of the stream; strong: the workload's reads are split over the ranks
    reads = w["reads"] // world if strong else w["reads"]          # per GPU
    total_reads = reads * world
    nbytes = reads * rb
    k, cov = w["k"], w["coverage"]
    bases_per_step = reads * READ_LEN                               # per GPU
    streamed = w["streamed"]
    first = rank * reads

    ctx = Context(local_rank, chunk_bytes=args.chunk_mib << 20)
    for kv in args.tune:
        name, v = kv.split("=")
        ctx.set_tuning(name, int(v))
    ctx.set_tuning("keep_stage_ms", 1)      # a step may be several count calls: the bench resets the stage timers itself
    GEN = 4_000_000

    def generate(d_ptr, lo, n):
        for off in range(0, n, GEN):
            m = min(GEN, n - off)
            ctx.synth_generate_device(params, lo + off, m, d_ptr + off * rb)

    batch_reads = min(args.batch_reads, reads)
    d_reads = ctx.device_alloc((batch_reads if streamed else reads) * rb)
    if not streamed:
        generate(d_reads, first, reads)      # device generator, not timed
    hint = bases_per_step // args.hint_div if args.hint_div > 0 else 0   # expected distinct k-mers; the table still grows (rehash) on demand beyond this

    sharded = world > 1 or args.force_sharded
    counter = None
    if sharded:
        from metacherchant_b200.dist import ShardedCounter
        if streamed and args.exchange != "merge":
            raise SystemExit("streamed workloads use the merge exchange")
        counter = ShardedCounter(ctx, k, world, rank, capacity_hint=hint, exchange=args.exchange)

    def timed_ms(fn):
        """device milliseconds of fn() on the library's stream (CUDA events) -- used where the generator runs between the calls"""
        ctx.event_record(2)
        r = fn()
        ctx.event_record(3)
        return r, ctx.event_elapsed_ms(2, 3)

    # ---- one step ------------------------------------------------------------------------------------------------------
    def step_device(keep_solid=False):
        """returns (global solid k-mers, stage ms dict, solid table or None, device ms of the counting calls if streamed)"""
        ctx.stage_reset()
        ms = 0.0
        if sharded:
            counter.step_begin()
            if streamed:
                counter.local_begin().expect_keys(reads * (READ_LEN - k + 1))
                for lo in range(0, reads, batch_reads):
                    n = min(batch_reads, reads - lo)
                    generate(d_reads, first + lo, n)
                    ms += timed_ms(lambda: counter.local_add_device(d_reads, n * rb, FMT_FASTQ))[1]
                _, t = timed_ms(counter.local_finish)
                ms += t
            else:
                counter.count_device(d_reads, nbytes, FMT_FASTQ, batch_bytes=args.chunk_mib << 20, record_bytes=rb)
            res, t = timed_ms(lambda: counter.step_finish(cov, keep_solid=keep_solid))
            ms += t
            return res[0], res[1], (res[2] if keep_solid else None), ms
        t = ctx.new_table(k, capacity_hint=hint)
        if streamed:
            t.expect_keys(reads * (READ_LEN - k + 1))
            for lo in range(0, reads, batch_reads):
                n = min(batch_reads, reads - lo)
                generate(d_reads, first + lo, n)
                ms += timed_ms(lambda: t.count_device(d_reads, n * rb, FMT_FASTQ))[1]
            solid, tf = timed_ms(lambda: t.filter(cov))
            ms += tf
        else:
            t.count_device(d_reads, nbytes, FMT_FASTQ)
            solid = t.filter(cov)
        n_solid = solid.size()
        stage = ctx.stage_ms()
        if not keep_solid:
            solid.free()
            solid = None
        t.free()
        return n_solid, stage, solid, ms

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    allocs0 = ctx.arena_stats()[1]
    sampler = ClockSampler(local_rank)
    if args.clocks == "smi":
        sampler.interval = None
    if not args.no_clocks:
        sampler.start()
    launches0 = ctx.launches
    ctx.event_record(0)
    t0 = time.perf_counter()
    kernel_ms = 0.0
    counted_ms = 0.0
    n_solid = 0
    stages = {}
    host_phases, step_walls = {}, []
    for _ in range(args.steps):
        tw = time.perf_counter()
        n_solid, stage, _, ms = step_device()
        step_walls.append(round(1000 * (time.perf_counter() - tw), 1))
        for name, v in ctx.host_ms().items():
            host_phases[name] = host_phases.get(name, 0.0) + v
        # the counting pipeline = extract (k-mers -> hashed keys) + scatter (keys -> bins) + count (bins -> table)
        kernel_ms += stage["extract"] + stage["scatter"] + stage["count"]
        counted_ms += ms
        for name, v in stage.items():
            stages[name] = stages.get(name, 0.0) + v
    ctx.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launches - launches0
    driver_allocs = ctx.arena_stats()[1] - allocs0
    clocks = sampler.stop()
    # resident workloads: the whole region (barrier to barrier); streamed ones: the counting calls only (the generator runs between them)
    t_ms = counted_ms if streamed else max(dev_ms, 0.0)
    if world > 1:
        tt = torch.tensor([t_ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)      # max over ranks
        t_ms, kernel_ms = tt.tolist()
    value = bases_per_step * world * args.steps / (t_ms / 1e3) / 1e9
    phases = {name: round(1000 * v, 2) for name, v in counter.phase_s.items()} if sharded else None   # last step, host view

    # ---- multi-GPU parity inside the bench: nothing lost or duplicated on the way to the owners -------------------------
    verify = None
    if sharded and not args.no_verify and args.exchange == "merge" and not streamed:
        counter.verify = True
        step_device()
        verify = counter.verify_exchange()
        counter.verify = False

    # ---- reads -> environments (the other half of BASELINE.json's metric), once, outside the timed region -----------
    env = None
    solid_pairs = None
    if not args.no_env:
        import shutil
        res = step_device(keep_solid=True)
        solid_local = res[2]
        barrier()
        t_g = time.perf_counter()
        if sharded:
            solid_full, per_rank = counter.gather_solid(solid_local)     # E3: all-gather of the solid (key, count16) set
            solid_local.free()
        else:
            solid_full, per_rank = solid_local, [res[0]]
        barrier()
        gather_s = time.perf_counter() - t_g
        out_root = args.keep_env_dir or tempfile.mkdtemp(prefix="mc_env_") if rank == 0 else None
        if world > 1:
            box = [out_root]
            dist.broadcast_object_list(box, src=0)
            out_root = box[0]
        try:
            fa = os.path.join(out_root, f"targets_{rank}.fasta")
            os.makedirs(out_root, exist_ok=True)
            with open(fa, "w") as f:
                f.write(s.targets_fasta())
            targets = ctx.load_targets(fa)
            names, seqs = [x[0] for x in targets], [x[1] for x in targets]
            barrier()
            t_env = time.perf_counter()
            kw = dict(min_cov=cov, max_kmers=args.maxkmers, both_dirs=w["bothdirs"])
            if sharded:
                found = counter.find_environments(solid_full, seqs, names, os.path.join(out_root, "out"), **kw)   # E4: targets split over the ranks
            else:
                found = ctx.find_environments(solid_full, seqs, names, out_dir=os.path.join(out_root, "out"), **kw)
            barrier()
            env_s = time.perf_counter() - t_env
            n_env_kmers, at_cap = 0, 0
            if rank == 0:
                for name in names:
                    g = os.path.join(out_root, "out", name, "env.txt")
                    if os.path.exists(g):
                        with open(g) as f:
                            lines = sum(1 for _ in f)
                        n_env_kmers += lines
                        at_cap += lines >= args.maxkmers      # one pass alone can hold --maxkmers k-mers beyond the seeds
            env = {"targets": len(targets), "found": int(sum(found)), "maxkmers": args.maxkmers, "bothdirs": w["bothdirs"], "env_kmers": n_env_kmers,
                   "targets_at_cap": at_cap, "env_ms": round(1000 * env_s, 1), "env_ms_per_target": round(1000 * env_s / max(1, len(targets)), 2),
                   "gather_ms": round(1000 * gather_s, 1) if sharded else 0.0, "solid_pairs_per_rank": per_rank,
                   "what": "BFS + unitigs + graph.gfa/env.txt/seqs.fasta/tsvs for all targets (mc_env_find, files written to a tmpfs dir); N > 1: "
                           "all-gather of the solid (key,count16) set, then targets block-distributed over the ranks",
                   "reads_to_env_s": round(t_ms / args.steps / 1e3 + gather_s + env_s, 3)}
            # keep the solid pairs for the CPU baseline of the same stage (rank 0, N = 1)
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                n_sp = solid_full.dump_device(1)
                kk = torch.empty(max(n_sp, 1), dtype=torch.int64, device="cuda")
                cc = torch.empty(max(n_sp, 1), dtype=torch.int16, device="cuda")
                solid_full.dump_device(1, kk.data_ptr(), cc.data_ptr(), n_sp)
                solid_pairs = (kk[:n_sp].cpu().numpy().view(np.uint64), cc[:n_sp].cpu().numpy(), fa if args.keep_env_dir else None, s.targets_fasta())
                del kk, cc
        finally:
            if not args.keep_env_dir:
                barrier()
                if rank == 0:
                    shutil.rmtree(out_root, ignore_errors=True)
            solid_full.free()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------
    e2e = None
    if not args.no_e2e and not streamed and (not sharded or args.exchange == "merge"):
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        ctx.d2h(host.numpy(), d_reads)
        ctx.device_free(d_reads)
        d_reads = None

        def step_host():
            ctx.stage_reset()
            if sharded:
                counter.step_begin()
                counter.local_begin()
                counter.local_add_host(host.data_ptr(), nbytes, FMT_FASTQ)
                counter.local_finish()
                return counter.step_finish(cov)[0]
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
        if world > 1:
            tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        e2e = {"value": round(bases_per_step * world * args.steps / e2e_s / 1e9, 4), "unit": "Gbases/s", "h2d_bytes_per_step": nbytes * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": round(1000 * e2e_s / args.steps, 2),
               "what": "pinned host FASTQ -> mc_count_bytes (chunked H2D overlapped with counting) -> [N > 1: merge exchange] -> coverage filter -> size read back"}
        del host

    # ---- roofline ------------------------------------------------------------------------------------
    peak, peak_src = load_peaks()
    alg_bytes = algorithmic_bytes_per_read(k, rb) * reads * args.steps   # per rank
    achieved = alg_bytes / (t_ms / 1e3) / 1e9 if t_ms > 0 else None
    achieved_pipe = alg_bytes / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None
    roofline = {"kernel": "whole step: parse + k_extract + k_bin_scatter + k_bin_scatter2 + k_bin_agg + table growth + coverage compaction (per rank)",
                "bound": "hbm", "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4) if achieved else None, "traffic": None, "peak_source": peak_src,
                "definition": "SURVEY 8d: algorithmic_bytes_per_read x reads per step / ms_per_step (the whole step, host gaps included)",
                "algorithmic_bytes_per_read": algorithmic_bytes_per_read(k, rb),
                "frac_pipeline": round(achieved_pipe / peak, 4) if achieved_pipe else None,
                "frac_pipeline_what": "the same bytes over the device time of k_extract + k_bin_scatter + k_bin_scatter2 + k_bin_agg only",
                "kernel_ms_per_step": round(kernel_ms / args.steps, 2),
                "stage_ms_per_step": {name: round(v / args.steps, 2) for name, v in stages.items()},
                "stage_sum_ms_per_step": round(sum(stages.values()) / args.steps, 2),
                "host_ms_per_step": {name: round(v / args.steps, 2) for name, v in host_phases.items()},
                "host_ms_what": "host wall time inside the same phases (launch + synchronisation waits included; count = the whole flush)",
                "step_wall_ms": step_walls}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1 and args.workload == "cfg2" and not args.reads:
        try:
            tj = json.load(open(prof))
            per_read = tj.get("step_dram_bytes_per_read") or tj.get("pipeline_dram_bytes_per_read")
            roofline["traffic"] = int(per_read * reads) if per_read else None
            roofline["traffic_note"] = tj.get("note", "ncu dram__bytes_read.sum + dram__bytes_write.sum of every launch of one cfg2 step (profiles/), NOT measured in this run")
        except Exception:
            pass
    if kernel_ms > 0:
        roofline["pipeline_share"] = {name: round(stages.get(name, 0.0) / kernel_ms, 3) for name in ("extract", "scatter", "count")}

    # ---- CPU baseline (rank 0, N = 1): the oracle's multi-threaded restatement on a bounded sample --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_reads = min(total_reads, args.cpu_sample_reads)
        data = s.generate(0, sample_reads, threads=cores)
        runs = [o.bench_count(data, FMT_FASTQ, k, cores) for _ in range(3)]
        secs = min(r[0] for r in runs)
        b = runs[0][1]
        cpu = {"value": round(b / secs / 1e9, 5), "unit": "Gbases/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_reads} reads of the workload ({data.size / 1e6:.0f} MB FASTQ), parse+count, C++ restatement of the Java reference, "
                         f"best of 3: {secs:.1f} s (median {sorted(r[0] for r in runs)[1]:.1f} s)"}
        del data
        if solid_pairs is not None and env is not None:
            # the same environment stage on the host cores: one OneSequenceCalculator per target on a pool of `cores` threads
            # (EnvironmentFinderMain.java:188,218-225) over a map holding the solid k-mers (every consumer only tests
            # get(key) >= coverage, OneSequenceCalculator.java:162,203: k-mers below the threshold behave like absent ones)
            import shutil
            kk, cc, _, fasta_text = solid_pairs
            ot = o.table(k, threads=cores)
            ot.add_pairs(kk, cc)
            d = tempfile.mkdtemp(prefix="mc_env_cpu_")
            try:
                fa = os.path.join(d, "targets.fasta")
                n_cpu_t = min(w["n_targets"], args.cpu_env_targets)
                open(fa, "w").write("".join(f">gene{t}\n{s.target(t)}\n" for t in range(n_cpu_t)))
                secs_env, done = o.bench_env(ot, fa, k, cov, args.maxkmers, None, w["bothdirs"], 1, cores, os.path.join(d, "out"))
                cpu["env_s"] = round(secs_env, 3)
                cpu["env_cores"] = cores
                cpu["env_targets"] = n_cpu_t
                cpu["env_targets_found"] = done
                cpu["env_s_per_target"] = round(secs_env / max(1, n_cpu_t), 3)
                cpu["env_sample"] = (f"all {len(kk)} solid k-mers of the workload (dumped from the GPU's coverage-filtered table) and the first {n_cpu_t} of {w['n_targets']} targets: "
                                     "BFS + unitigs + every output file, one calculator per target on a thread pool")
                cpu["reads_to_env_s_extrapolated"] = round(total_reads * READ_LEN / (cpu["value"] * 1e9) + secs_env * w["n_targets"] / max(1, n_cpu_t), 1)
            finally:
                shutil.rmtree(d, ignore_errors=True)
            del ot

    if rank == 0:
        timing = ("CUDA events around each batch's counting call + the merge/filter calls; the on-device read generator runs between them and is excluded"
                  if streamed else "barrier + synchronize on both sides of the K steps, CUDA events on the library's stream, max over ranks")
        line = {"metric": "kmer_count_throughput", "value": round(value, 4), "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(t_ms / args.steps, 2), "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
                "dtype": "int64", "data": "synthetic",
                "config": {"workload": w["name"], "description": w["desc"], "k": k, "coverage": cov, "reads_per_gpu_per_step": reads, "reads_total_per_step": total_reads,
                           "fastq_bytes_per_gpu_per_step": nbytes, "l2": "inputs (>= 0.3 GB per step) larger than the 126 MB L2; no explicit flush",
                           "chunk_mib": args.chunk_mib, "timing": timing, "batch_reads": batch_reads if streamed else None,
                           "table": (f"pre-sized for bases/{args.hint_div} distinct k-mers, grows by rehash beyond" if args.hint_div > 0 else
                                     "no capacity hint: sized by the library from a counted 1/64 sample of the hash space (load 0.5), grows by rehash beyond"),
                           "solid_kmers": int(n_solid), "wall_s": round(wall, 3), "exchange": args.exchange if sharded else None},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "fused_dump": (counter.fused_last if sharded else None), "driver_allocs_in_timed_region": int(driver_allocs),
                "roofline": roofline, "cpu_baseline": cpu, "env": env, "phases_ms_last_step": phases, "verify": verify}
        print(json.dumps(line), flush=True)
    if sharded:
        counter.close_exchange()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
