# per-kernel durations of one FIXED-base MSM at 2^24 with the pair tree (usage: TUNE=... bash scripts/ncu_pairs.sh tag)
tag=${1:-pairs}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_pair|k_prod|k_inv|k_accumulate|k_digits' -c 32 --csv --log-file gpurun_out/s2_ncu_$tag.csv python scripts/gpu_pairs_probe.py 24 0 4 > gpurun_out/s2_ncu_$tag.log 2>&1; tail -1 gpurun_out/s2_ncu_$tag.log | cut -c1-200
