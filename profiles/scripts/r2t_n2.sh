# round 2, capture "t": fused flush + owner dump -- parity worker and bench at N ranks
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_parity_worker.py > gpurun_out/r2t_dist_parity_n$N.log 2>&1; grep -E "DIST_|Error|error" gpurun_out/r2t_dist_parity_n$N.log | head
for fuse in 1 0; do
MC_FUSE_DUMP=$fuse timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 4 --warmup 2 --no-env --no-e2e > gpurun_out/r2t_bench_n${N}_fuse$fuse.log 2> gpurun_out/r2t_bench_n${N}_fuse$fuse.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench_n${N}_fuse$fuse.log') if l.startswith('{')][-1]); r=d['roofline']; print('N=$N fuse=$fuse', d['value'], d['ms_per_step'], 'fused', d['fused_dump'], d['phases_ms_last_step'], d['verify'] and d['verify']['ok'])"
tail -3 gpurun_out/r2t_bench_n${N}_fuse$fuse.err | grep -v OMP | grep -v "\*\*\*"
done
