set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 300 python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 2500 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
timeout 200 python bench.py --streams 1 --no-single --no-cpu-baseline > gpurun_out/bench_r1_b_s1.json 2>> gpurun_out/bench_r1_b.err; tail -c 1200 gpurun_out/bench_r1_b_s1.json
timeout 200 python bench.py --streams 15 --no-single --no-cpu-baseline > gpurun_out/bench_r1_b_s15.json 2>> gpurun_out/bench_r1_b.err; tail -c 1200 gpurun_out/bench_r1_b_s15.json
