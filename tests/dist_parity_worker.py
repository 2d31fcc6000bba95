"""Run under torchrun with N ranks (N = 2, 4, 8) on one box: checks that the hash-sharded table (union of the
shards after the NCCL all-to-all) equals the oracle's single table bit-exactly, whatever N.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_ffi
    from metacherchant_b200 import Context
    from metacherchant_b200.dist import ShardedCounter
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    o = oracle_ffi.load()
    s = o.synth(genome_unit=20_000, n_targets=2)
    total_reads, k = 120_000, 31
    per = total_reads // world
    ctx = Context(local)
    rb = s.record_bytes
    d = ctx.device_alloc(per * rb)
    ctx.synth_generate_device(s.params_bytes(), rank * per, per, d)
    # both exchanges must give the same shards: NCCL all-to-all and the peer-memory scatter (MC_DIST_EXCHANGE picks one to dump)
    tables = {}
    env_ok, verify = True, None
    import tempfile
    shared = [tempfile.mkdtemp(prefix="mc_dist_env_") if rank == 0 else None]
    dist.broadcast_object_list(shared, src=0)
    shared = shared[0]
    for mode in ("nccl", "peer", "merge"):
        counter = ShardedCounter(ctx, k, world, rank, exchange=mode)
        counter.verify = mode == "merge"
        counter.new_table()
        counter.count_device(d, per * rb, oracle_ffi.FMT_FASTQ, batch_bytes=3_000_000, record_bytes=rb)   # several batches
        tables[mode] = counter.table.dump()
        if mode == "merge":
            # once more: the exchange regions exist now, so the owner dump rides on the local table's last flush (mc_xchg_flush_and_dump;
            # "agg" forces the aggregated flush this small input would not take by itself)
            ctx.set_tuning("agg", 1)
            counter.fuse_dump = True
            counter.new_table()
            counter.count_device(d, per * rb, oracle_ffi.FMT_FASTQ, batch_bytes=3_000_000, record_bytes=rb)
            again = counter.table.dump()
            counter.fuse_dump = False
            ctx.set_tuning("agg", -1)
            fused_ok = np.array_equal(again[0], tables[mode][0]) and np.array_equal(again[1], tables[mode][1])
            fl = torch.tensor([1 if (fused_ok and counter.fused_last == 1) else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(fl, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"DIST_FUSED_DUMP world={world} fused={counter.fused_last} {'OK' if int(fl.item()) == 1 else 'MISMATCH'}", flush=True)
            env_ok = env_ok and int(fl.item()) == 1
        if mode == "merge":
            # bench.py --verify's check: the digests of the local tables and of the owner shards add up to the same totals
            verify = counter.verify_exchange()
            # E3 + E4: local compaction, all-gather of the solid set, targets split over the ranks
            cov = 3
            solid = counter.table.filter(cov)
            full, per_rank = counter.gather_solid(solid)
            solid.free()
            fa = os.path.join(shared, f"genes_{rank}.fasta")
            open(fa, "w").write(s.targets_fasta())
            targets = ctx.load_targets(fa)
            found = counter.find_environments(full, [t[1] for t in targets], [t[0] for t in targets], os.path.join(shared, "gpu"),
                                              min_cov=cov, max_kmers=3000, both_dirs=False)
            full.free()
            dist.barrier()
            if rank == 0:
                ref_t = o.table(k, threads=8).count_bytes(s.generate(0, per * world), oracle_ffi.FMT_FASTQ)
                n_ref = ref_t.env_find(fa, cov, 3000, None, False, 1, os.path.join(shared, "orc"))
                env_ok = n_ref == sum(found) == len(targets)
                for name, _ in targets:
                    for f in ("env.txt", "graph.txt", "graph.gfa", "seqs.fasta", "tsvs/nodes.tsv", "tsvs/edges.tsv"):
                        a = open(os.path.join(shared, "gpu", name, f), "rb").read()
                        b = open(os.path.join(shared, "orc", name, f), "rb").read()
                        env_ok = env_ok and a == b
                print(f"DIST_ENV_PARITY world={world} targets={len(targets)} found={sum(found)} solid_pairs_per_rank={per_rank} "
                      f"verify={verify['ok']} digest={verify['digest']} {'OK' if env_ok else 'MISMATCH'}", flush=True)
        stats = counter.table.stats()
        tot = torch.tensor([stats["reads"], stats["kmers"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert int(tot[0].item()) == per * world, (mode, stats)
        counter.close_exchange()
    same = all(np.array_equal(tables["nccl"][i], tables[m][i]) for m in ("peer", "merge") for i in (0, 1))
    keys, counts = tables["peer"]
    # gather the shards on rank 0
    n = torch.tensor([keys.size], dtype=torch.int64, device="cuda")
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(x.item() for x in sizes))
    kbuf = torch.zeros(m, dtype=torch.int64, device="cuda")
    cbuf = torch.zeros(m, dtype=torch.int32, device="cuda")   # NCCL has no int16
    kbuf[: keys.size] = torch.from_numpy(keys.view(np.int64)).cuda()
    cbuf[: keys.size] = torch.from_numpy(counts.astype(np.int32)).cuda()
    ks = [torch.zeros(m, dtype=torch.int64, device="cuda") for _ in range(world)]
    cs = [torch.zeros(m, dtype=torch.int32, device="cuda") for _ in range(world)]
    dist.all_gather(ks, kbuf)
    dist.all_gather(cs, cbuf)
    ok = True
    if rank == 0:
        allk = np.concatenate([ks[r][: int(sizes[r].item())].cpu().numpy().view(np.uint64) for r in range(world)])
        allc = np.concatenate([cs[r][: int(sizes[r].item())].cpu().numpy().astype(np.int16) for r in range(world)])
        order = np.argsort(allk, kind="stable")
        allk, allc = allk[order], allc[order]
        ref = o.table(k, threads=8).count_bytes(s.generate(0, per * world), oracle_ffi.FMT_FASTQ)
        rk, rc = ref.dump()
        ok = np.array_equal(allk, rk) and np.array_equal(allc, rc) and len(np.unique(allk)) == allk.size and same and env_ok
        print(f"DIST_PARITY world={world} distinct={allk.size} oracle={rk.size} exchanged_bytes_rank0={counter.bytes_exchanged} {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
