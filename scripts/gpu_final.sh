# Final evidence run of the round: GPU tests, smoke, default bench (both arms), C4, launch list, captures.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -3 gpurun_out/bench_r1_final.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_r1_final_ref.json 2>> gpurun_out/bench_r1_final.err
timeout 400 python bench.py --config C4 --streams 1 --steps 60 --warmup 5 --no-single --no-cpu-baseline > gpurun_out/bench_r1_c4.json 2> gpurun_out/bench_r1_c4.err; tail -2 gpurun_out/bench_r1_c4.err
for S in 12 16; do timeout 300 python bench.py --streams $S --steps 100 --no-single --no-cpu-baseline > gpurun_out/bench_r1_s$S.json 2>/dev/null; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 10 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -o gpurun_out/prof_grid_r1_final python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -o gpurun_out/prof_grid_r1_c4 python bench.py --config C4 --streams 1 --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_epipolar_search -s 3 -c 1 -o gpurun_out/prof_epi_r1_final2 python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1



ls -la gpurun_out | tail -14
