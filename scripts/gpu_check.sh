# Quick health check of the committed state: GPU tests, smoke(), fb_update stage timings, default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/profile_update.py 16; python scripts/profile_update.py 8
timeout 500 python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; tail -2 gpurun_out/bench_check.err
python -c "
import json; d=json.load(open('gpurun_out/bench_check.json'))
print('value %.0f e2e %.0f e2e_sync %.0f solver_us %.1f frac %.3f single %.0f/%.0f cpu %.0f'%(d['value'],d['e2e']['value'],d['e2e_sync']['value'],d['roofline']['launch_us'],d['roofline']['frac'],d['single_stream']['value'],d['single_stream']['e2e'],d['cpu_baseline']['value']))"
