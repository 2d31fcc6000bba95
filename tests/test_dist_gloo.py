"""N>1 host path on CPU: the key exchange of the hash-sharded counting step (metacherchant_b200/dist.py) under
gloo, world_size 2 — ownership by the top hash bits, uneven group sizes, empty groups, and the exactness
property the GPU path relies on (the union of the shards equals the single table)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metacherchant_b200.dist import exchange_keys, owner_of, plan_batches, shard_bits_for


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bits = shard_bits_for(world)
        rng = np.random.default_rng(100 + rank)
        for batch, n in enumerate((5000, 0 if rank == 0 else 37, 1)):
            keys = rng.integers(0, 1 << 62, size=n, dtype=np.int64)
            if n:
                keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3]      # duplicates across and inside ranks
            owners = np.array([owner_of(int(h), bits) for h in keys], dtype=np.int64)
            order = np.argsort(owners, kind="stable")
            send = torch.from_numpy(keys[order].copy())
            counts = [int((owners == r).sum()) for r in range(world)]
            recv, recv_counts = exchange_keys(send, counts)
            got = recv.numpy()
            assert got.size == sum(recv_counts)
            assert all(owner_of(int(h), bits) == rank for h in got)       # everything I received is mine
            np.save(os.path.join(out_dir, f"sent_{rank}_{batch}.npy"), keys)
            np.save(os.path.join(out_dir, f"recv_{rank}_{batch}.npy"), got)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_exchange_keys_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for batch in range(3):
        sent = np.concatenate([np.load(tmp_path / f"sent_{r}_{batch}.npy") for r in range(world)])
        recv = np.concatenate([np.load(tmp_path / f"recv_{r}_{batch}.npy") for r in range(world)])
        # nothing lost, nothing duplicated: the multiset of keys is preserved, so per-key counts are exact sums
        assert np.array_equal(np.sort(sent), np.sort(recv))


def test_owner_and_batch_planning():
    assert shard_bits_for(1) == 0 and shard_bits_for(2) == 1 and shard_bits_for(8) == 3
    with pytest.raises(ValueError):
        shard_bits_for(3)
    assert owner_of((1 << 62) - 1, 3) == 7 and owner_of(0, 3) == 0 and owner_of(1 << 61, 1) == 1 and owner_of(123, 0) == 0
    # fixed-size records: slices are whole records and cover the buffer exactly
    b = plan_batches(315 * 1000, 100_000, 315)
    assert sum(n for _, n in b) == 315 * 1000 and all(off % 315 == 0 and n % 315 == 0 for off, n in b)
    assert plan_batches(0, 100, 7) == []
    assert plan_batches(10, 4, None) == [(0, 4), (4, 4), (8, 2)]
    assert plan_batches(315 * 3, 10, 315) == [(0, 315), (315, 315), (630, 315)]   # a batch is never smaller than a record


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from metacherchant_b200.dist import gather_pairs
        rng = np.random.default_rng(7 + rank)
        for rnd, n in enumerate((1000 + 300 * rank, 0 if rank == 1 else 5, 0)):   # uneven, one empty rank, all empty
            keys = torch.from_numpy(rng.integers(0, 1 << 62, size=max(n, 1), dtype=np.int64))
            counts = torch.from_numpy(rng.integers(1, 32768, size=max(n, 1), dtype=np.int16))
            parts = gather_pairs(keys, counts, n)
            assert len(parts) == world
            assert torch.equal(parts[rank][0], keys[:n]) and torch.equal(parts[rank][1], counts[:n])
            np.save(os.path.join(out_dir, f"mine_{rank}_{rnd}.npy"), keys[:n].numpy())
            np.save(os.path.join(out_dir, f"all_{rank}_{rnd}.npy"), torch.cat([p[0] for p in parts]).numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_pairs_world2(tmp_path):
    """E3: after the all-gather every rank holds the pairs of ALL ranks, in rank order, also when ranks
    contribute different amounts or nothing."""
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rnd in range(3):
        want = np.concatenate([np.load(tmp_path / f"mine_{r}_{rnd}.npy") for r in range(world)])
        for r in range(world):
            assert np.array_equal(np.load(tmp_path / f"all_{r}_{rnd}.npy"), want)


def test_target_blocks_cover_every_target_once():
    """E4: targets block-distributed over the ranks (EnvironmentFinderMain.java:218-225 runs them independently)."""
    from metacherchant_b200.dist import target_block
    for n_targets in (0, 1, 3, 20, 500):
        for world in (1, 2, 4, 8):
            blocks = [target_block(n_targets, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n_targets
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
