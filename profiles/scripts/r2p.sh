# round 2, capture "p": (1) where the "grow" time of the sharded step goes (MC_TRACE), (2) cfg2 IN FULL against the oracle's digest (once per round)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 1 --warmup 1"
MC_TRACE=1 $B --force-sharded > gpurun_out/r2p_trace.log 2> gpurun_out/r2p_trace.err; grep "table grown" gpurun_out/r2p_trace.err | tail -8
MC_RUN_SLOW=1 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "cfg2_in_full" 2>&1 | grep -E "CFG2_FULL_PARITY|passed|failed" > gpurun_out/r2p_cfg2_full_parity.log; cat gpurun_out/r2p_cfg2_full_parity.log | cut -c1-900
