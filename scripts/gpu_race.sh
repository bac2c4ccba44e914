timeout 600 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_golden.py -q -x 2>&1 | tail -2
timeout 500 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_gpu_nltgv2.py -q -x -k "short_solves or small_graph_parity or partition_and_transport" 2>&1 | grep -v "^=========     at\|^=========     by\|Saved host" | tail -12
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_nltgv2.py -q -x -k "short_solves or small_graph_parity or partition_and_transport or cta_shapes" 2>&1 | tail -3
for S in 1 8; do timeout 200 python bench.py --streams $S --steps 100 --no-single --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('S=$S value %.0f e2e %.0f solver_us %.1f frac %.2f'%(d['value'],d['e2e']['value'],d['roofline']['launch_us'],d['roofline']['frac']))"; done
