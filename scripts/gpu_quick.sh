set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err; tail -3 gpurun_out/bench_r1_h.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1_h.json"))
print("value %.0f e2e(pipe) %.0f e2e_sync %.0f solver_us %.1f frac %.2f cpu %.0f"%(d["value"],d["e2e"]["value"],d["e2e_sync"]["value"],d["roofline"]["launch_us"],d["roofline"]["frac"],d["cpu_baseline"]["value"]))
print(d["single_stream"])
PY
timeout 200 python bench.py --streams 15 --no-single --no-cpu-baseline > gpurun_out/bench_r1_h_s15.json 2>> gpurun_out/bench_r1_h.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1_h_s15.json"))
print("S15 value %.0f e2e(pipe) %.0f e2e_sync %.0f solver_us %.1f frac %.2f"%(d["value"],d["e2e"]["value"],d["e2e_sync"]["value"],d["roofline"]["launch_us"],d["roofline"]["frac"]))
PY
