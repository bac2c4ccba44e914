# Round-1 evidence run: GPU tests, default bench (both arms), ncu launch list + full captures.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -3 gpurun_out/bench_r1_final.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_r1_final_ref.json 2>> gpurun_out/bench_r1_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 10 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_cluster -s 3 -c 1 -o gpurun_out/prof_cluster_r1_final python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_epipolar_search -s 3 -c 1 -o gpurun_out/prof_epi_r1_final python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
