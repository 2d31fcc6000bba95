# ncu captures of the binned counting pipeline (round 1, capture "c"); 8 M reads so that ~40 replays per kernel stay short
mkdir -p gpurun_out
B="python bench.py --reads 8000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv $B > gpurun_out/r1c_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extract|k_bin_scatter' -s 6 -c 2 -o gpurun_out/r1c_extract_scatter $B > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bin_count -s 12 -c 1 -o gpurun_out/r1c_bin_count $B > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_count_terms|k_write_terms' -s 6 -c 2 -o gpurun_out/r1c_parse $B > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rehash|k_dump' -s 2 -c 2 -o gpurun_out/r1c_filter $B > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
