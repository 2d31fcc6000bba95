run() { timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'], d['e2e'])"; }
for k in 1 2 3 4; do run --tune host_batch_chunks=$k; done
