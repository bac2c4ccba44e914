mkdir -p gpurun_out
for mode in spec nospec; do
  if [ $mode = nospec ]; then export FB_NO_SPEC_MAP=1; else unset FB_NO_SPEC_MAP; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_spec_$mode.csv --launch-skip 600 --launch-count 300 python scripts/profile_update.py 8 0 50 getter > /dev/null 2>&1
  echo "== $mode"; python scripts/launch_table.py gpurun_out/r2_spec_$mode.csv | grep -E "raster|validity|all"
  timeout 100 python scripts/profile_update.py 8 0 50 getter | cut -c1-130
done
