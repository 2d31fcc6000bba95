"""world = 1 run of the sharded counting path with per-phase wall times (profiling aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import oracle_ffi
from metacherchant_b200 import Context
from metacherchant_b200.dist import ShardedCounter
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
o = oracle_ffi.load()
s = o.synth(n_genomes=10, genome_unit=1_000_000, read_len=150, n_targets=20, target_len=900, repeat_len=0, err_ppm=5000, n_ppm=100, bangq_ppm=100, fasta=False)
ctx = Context(0, chunk_bytes=1 << 30)
rb = s.record_bytes
d = ctx.device_alloc(reads * rb)
for off in range(0, reads, 4_000_000):
    n = min(4_000_000, reads - off)
    ctx.synth_generate_device(s.params_bytes(), off, n, d + off * rb)
c = ShardedCounter(ctx, 31, 1, 0, capacity_hint=reads * 150 // 4)
for it in range(3):
    torch.cuda.synchronize(); ctx.sync(); t0 = time.perf_counter()
    ctx.stage_reset()
    c.new_table()
    t1 = time.perf_counter()
    c.count_device(d, reads * rb, 1, batch_bytes=1 << 30, record_bytes=rb)
    ctx.sync(); t2 = time.perf_counter()
    solid = c.table.filter(5)
    n = solid.size(); ctx.sync(); t3 = time.perf_counter()
    print(f"iter {it}: new_table {1e3*(t1-t0):.1f} ms, count_device {1e3*(t2-t1):.1f} ms, filter(+flush) {1e3*(t3-t2):.1f} ms, solid {n}, stages {ctx.stage_ms()}")
    solid.free()
