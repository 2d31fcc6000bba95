# session 3: sweep of the bin geometry (first-level bits, aggregation unit size) at cfg2
mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-env "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*', d['ms_per_step'], r['stage_ms_per_step'])"; }
run
run --tune agg_unit=65536
run --tune agg_unit=16384
run --tune part_bits=8
run --tune part_bits=10
run --tune part_bits=10 --tune agg_unit=65536
run --tune part_bits=8 --tune agg_unit=65536
