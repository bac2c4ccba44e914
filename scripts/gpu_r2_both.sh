timeout 1200 python -m pytest tests/test_gpu_hotpath_step.py tests/test_gpu_delaunay.py tests/test_gpu_update.py tests/test_gpu_golden.py -q -x 2>&1 | tail -4
bash scripts/gpu_r2_launches.sh 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 3 --no-update --no-c4 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f (%.1f us) e2e %.0f (%.1f us) sync %.0f solver_us %.1f single %s' % (d['value'], 1e3*d['ms_per_step'], d['e2e']['value'], 1e3*d['e2e']['ms_per_step'], d['e2e_sync']['value'], d['roofline']['launch_us'], d['single_stream']))"
