set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -c 600 gpurun_out/bench_r1_g.json; tail -3 gpurun_out/bench_r1_g.err
./tests/cpp/_build/flame_shim_demo
