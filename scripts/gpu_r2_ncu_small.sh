# ncu --set full of the single-CTA graph-build kernels (source-level sampling: where a one-CTA kernel waits)
mkdir -p gpurun_out
for k in k_ds_prepare k_tile_assign k_ds_stars; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 25 -c 1 -f -o gpurun_out/r2_prof_$k python scripts/profile_update.py 8 0 > gpurun_out/r2_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
