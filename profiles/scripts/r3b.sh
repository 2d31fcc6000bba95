mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
MC_TRACE=1 $B --steps 2 --warmup 2 > gpurun_out/r3b_trace.log 2> gpurun_out/r3b_trace.err
grep -n "arena\|table grown\|step" gpurun_out/r3b_trace.err | tail -40
