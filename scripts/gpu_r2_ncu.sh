# Round 2 ncu --set full captures: the dominant kernel of each measured workload (one launch each).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline --no-update --no-c4"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -f -o gpurun_out/r2_prof_grid_c2x8 $B > gpurun_out/r2_ncu_grid.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -f -o gpurun_out/r2_prof_grid_c4 $B --config C4 --streams 1 > gpurun_out/r2_ncu_grid_c4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_tile -s 25 -c 1 -f -o gpurun_out/r2_prof_tile python scripts/profile_update.py 8 0 > gpurun_out/r2_ncu_tile.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ds_stars -s 25 -c 1 -f -o gpurun_out/r2_prof_stars python scripts/profile_update.py 8 0 > gpurun_out/r2_ncu_stars.log 2>&1
ls -la gpurun_out/*.ncu-rep
