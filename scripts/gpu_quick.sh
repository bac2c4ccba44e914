set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for S in 1 8 15; do
timeout 200 python bench.py --streams $S --no-single --no-cpu-baseline > gpurun_out/bench_r1_f_s$S.json 2>> gpurun_out/bench_r1_f.err
done
FB_CLUSTER_MIN=16 timeout 200 python bench.py --streams 1 --no-single --no-cpu-baseline > gpurun_out/bench_r1_f_c16_s1.json 2>> gpurun_out/bench_r1_f.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_f.csv python bench.py --steps 10 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ls gpurun_out
