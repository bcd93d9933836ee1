# ncu --set full of the first forward / backward launches of the pair tree (pass 0 and pass 1) at 2^24, FIXED base
ncu --set full --clock-control none --import-source on -k regex:'k_pair_(fwd|bwd)' -c 4 -o gpurun_out/r02_pair_tree_full -f python scripts/gpu_pairs_probe.py 24 0 4 > gpurun_out/ncu_pairs_full.log 2>&1
tail -2 gpurun_out/ncu_pairs_full.log | cut -c1-200
ncu -i gpurun_out/r02_pair_tree_full.ncu-rep --page details --csv > gpurun_out/r02_pair_tree_full_details.csv 2>/dev/null
ls -la gpurun_out/r02_pair_tree_full*
