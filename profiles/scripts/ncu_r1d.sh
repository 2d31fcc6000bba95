# ncu captures of the counting pipeline with two-level bins + shared-memory aggregation (round 1, capture "d")
# run on the GPU box from the repo root:  bash profiles/scripts/ncu_r1d.sh
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-env --no-cpu-baseline"
# 1. launch list of one full cfg2 step (50 M reads): duration + DRAM bytes of every launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1d_launches.csv $B > gpurun_out/r1d_launches.log 2>&1
# 2. full captures (source imported) of the four pipeline kernels, one launch each, from the middle of the step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extract|k_bin_scatter$' -s 12 -c 2 -o gpurun_out/r1d_extract_scatter $B > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter2|k_bin_agg' -s 7 -c 2 -o gpurun_out/r1d_scatter2_agg $B > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
