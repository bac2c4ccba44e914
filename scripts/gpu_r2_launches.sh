# Per-kernel durations of the fb_update pipeline (ncu launch list; cold-cache, serialised times).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_update_launches.csv \
  --launch-skip 500 --launch-count 300 python scripts/profile_update.py 8 0 > gpurun_out/r2_update_ncu.log 2>&1
tail -3 gpurun_out/r2_update_ncu.log
python scripts/launch_table.py gpurun_out/r2_update_launches.csv | tee gpurun_out/r2_update_launch_table.md
