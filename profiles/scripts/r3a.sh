# session 3: full GPU suite + smoke + default bench on the current tree
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3a_pytest.log; tail -3 gpurun_out/r3a_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
python bench.py > gpurun_out/r3a_bench.log 2> gpurun_out/r3a_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3a_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], 'e2e', d['e2e']['value'], 'env', d['env']['env_ms'], d['env']['reads_to_env_s'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('env_s'), d['clocks'])"
tail -2 gpurun_out/r3a_bench.err
