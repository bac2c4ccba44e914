# Minimal post-change sanity: pipelined hot path + golden vectors + a short bench.
timeout 200 python -m pytest tests/test_gpu_hotpath_step.py tests/test_gpu_golden.py -q -x 2>&1 | tail -1
timeout 150 python bench.py --steps 100 --no-single --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f solver_us %.1f frac %.3f'%(d['value'],d['e2e']['value'],d['roofline']['launch_us'],d['roofline']['frac']))"
