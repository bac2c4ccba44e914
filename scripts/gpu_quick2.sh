timeout 300 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_golden.py tests/test_gpu_hotpath_step.py -q -x 2>&1 | tail -1
for S in 1 4 8 12 15; do
timeout 300 python bench.py --streams $S --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('S=$S C=%s value %.0f (%.3f ms) e2e(pipe) %.0f e2e_sync %.0f solver_us %.1f frac %.2f'%(d['config']['sms_per_stream'],d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e_sync']['value'],d['roofline']['launch_us'],d['roofline']['frac']))"
tail -2 /tmp/err.txt
done
