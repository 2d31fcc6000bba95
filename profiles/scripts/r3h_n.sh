# session 3: parity worker and bench at N ranks on the current tree
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_parity_worker.py > gpurun_out/r3h_dist_parity_n$N.log 2>&1; grep -E "DIST_|Error|error" gpurun_out/r3h_dist_parity_n$N.log | head
MC_DIST_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3h_bench_n$N.log 2> gpurun_out/r3h_bench_n$N.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3h_bench_n$N.log') if l.startswith('{')][-1]); r=d['roofline']; print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms_last_step'], d['verify'] and d['verify']['ok'], 'env', d['env']['env_ms'], d['env']['gather_ms'], d['env']['reads_to_env_s'])"
grep "gather_solid" gpurun_out/r3h_bench_n$N.err | head -8; true
