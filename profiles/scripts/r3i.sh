# session 3, capture "i": N ranks -- parity worker + bench (cfg2 weak scaling) [+ cfg5 strong scaling when asked]
mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_parity_worker.py > gpurun_out/r3i_dist_parity_n$N.log 2>&1; grep -E "DIST_|Error|error" gpurun_out/r3i_dist_parity_n$N.log | head
MC_DIST_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3i_bench_n$N.log 2> gpurun_out/r3i_bench_n$N.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3i_bench_n$N.log') if l.startswith('{')][-1]); r=d['roofline']; print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], d['phases_ms_last_step'], r['stage_ms_per_step']); print('  env', d['env']); print('  verify', d['verify'])"
tail -3 gpurun_out/r3i_bench_n$N.err
if [ "$2" = "cfg5" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload cfg5 --steps 1 --warmup 1 --no-env > gpurun_out/r3i_bench_cfg5_n$N.log 2> gpurun_out/r3i_bench_cfg5_n$N.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3i_bench_cfg5_n$N.log') if l.startswith('{')][-1]); r=d['roofline']; print('cfg5 N=$N', d['value'], d['ms_per_step'], d['phases_ms_last_step'], r['stage_ms_per_step'], d['config']['solid_kmers'])"
tail -3 gpurun_out/r3i_bench_cfg5_n$N.err
fi
