run() { timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'], d['e2e'])"; }
run --no-e2e --reads 16000000
run --no-e2e --reads 16000000 --tune extract_parts=4
run --no-e2e
