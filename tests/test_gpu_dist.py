"""GPU: the N-rank sharded table equals the oracle's (runs only when the box has >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_counting_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "dist_parity_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_PARITY" in r.stdout and "DIST_ENV_PARITY" in r.stdout and "MISMATCH" not in r.stdout
    assert r.stdout.count(" OK") >= 2, r.stdout[-2000:]


def test_single_rank_sharded_path_matches_direct_path(oracle):
    """world = 1 exercises extract -> owner grouping -> insert_hashed on one GPU and must equal the fused kernel."""
    import numpy as np
    from metacherchant_b200 import Context
    from metacherchant_b200.dist import ShardedCounter
    import oracle_ffi
    s = oracle.synth(genome_unit=20_000)
    data = s.generate(0, 30_000)
    ctx = Context(0)
    try:
        d = ctx.device_alloc(data.size)
        ctx.h2d(d, data)
        c = ShardedCounter(ctx, 31, 1, 0)
        c.new_table()
        c.count_device(d, data.size, oracle_ffi.FMT_FASTQ, batch_bytes=2_000_000, record_bytes=s.record_bytes)
        a = c.table.dump()
        b = ctx.new_table(31).count_device(d, data.size, oracle_ffi.FMT_FASTQ).dump()
        ref = oracle.table(31, threads=4).count_bytes(data, oracle_ffi.FMT_FASTQ)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[0], ref.dump()[0]) and np.array_equal(a[1], ref.dump()[1])
        assert c.table.stats() == ref.stats()
    finally:
        ctx.close()


def test_two_devices_in_one_process(oracle):
    """One process, two contexts on two GPUs (a JVM that opens two devices): the kernels' shared-memory opt-in is per DEVICE, so the
    second context must configure its own (round 1 guarded it with a process-wide flag and the second device's launches failed)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from metacherchant_b200 import Context
    import oracle_ffi
    s = oracle.synth(genome_unit=10_000)
    data = s.generate(0, 40_000)
    ref = oracle.table(31, threads=4).count_bytes(data, oracle_ffi.FMT_FASTQ).dump()
    ctxs = [Context(0), Context(1)]
    try:
        for c in ctxs:
            c.set_tuning("agg", 1)          # the full-size path: scatter2 + aggregation kernels with > 48 KB of dynamic shared memory
            got = c.new_table(31).count_bytes(data, oracle_ffi.FMT_FASTQ).dump()
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    finally:
        for c in ctxs:
            c.close()
