mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ds_stars -s 25 -c 1 -f -o gpurun_out/r2_prof_stars_v2 python scripts/profile_update.py 8 0 > gpurun_out/r2_ncu_stars_v2.log 2>&1
ls -la gpurun_out/r2_prof_stars_v2.ncu-rep
