# Tile solver: prologue vs per-iteration cost (ncu launch list at 1 / 50 / 100 iterations per frame).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nltgv2.py -q -x -k "variant or tile or plan_free" 2>&1 | tail -3
for it in 1 50 100; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_tile_it$it.csv \
    --launch-skip 20 --launch-count 20 -k regex:k_nltgv2_tile python scripts/profile_update.py 8 0 $it > /dev/null 2>&1
  echo "iters $it"; python scripts/launch_table.py gpurun_out/r2_tile_it$it.csv | tail -2
done
timeout 120 python scripts/profile_update.py 8 0
