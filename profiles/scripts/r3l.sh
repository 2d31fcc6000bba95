mkdir -p gpurun_out
MC_TRACE=1 python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r3l_trace.log 2> gpurun_out/r3l_trace.err
grep "mc trace" gpurun_out/r3l_trace.err | grep -v "agg group\|arena\|table grown\|sk expand" | tail -30
