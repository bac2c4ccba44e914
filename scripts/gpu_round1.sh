set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
python bench.py --variant 1 --no-single --no-cpu-baseline > gpurun_out/bench_r1_streaming.json 2>> gpurun_out/bench_r1_a.err; tail -c 1500 gpurun_out/bench_r1_streaming.json
python bench.py --streams 1 --no-single --no-cpu-baseline > gpurun_out/bench_r1_s1.json 2>> gpurun_out/bench_r1_a.err; tail -c 1500 gpurun_out/bench_r1_s1.json
python bench.py --streams 16 --no-single --no-cpu-baseline > gpurun_out/bench_r1_s16.json 2>> gpurun_out/bench_r1_a.err; tail -c 1500 gpurun_out/bench_r1_s16.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1_a.err; tail -c 1500 gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_cluster -s 3 -c 2 -o gpurun_out/prof_cluster_r1 python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:k_epipolar_search -s 3 -c 2 -o gpurun_out/prof_epi_r1 python bench.py --steps 6 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
