timeout 300 python -m pytest tests/test_gpu_hotpath_step.py -q -x 2>&1 | tail -1
for S in 1 8 15; do
timeout 300 python bench.py --streams $S --no-single --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('S=$S value %.0f (%.3f ms) e2e(pipe) %.0f e2e_sync %.0f solver_us %.1f frac %.2f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e_sync']['value'],d['roofline']['launch_us'],d['roofline']['frac'],d['gpu_launches']))"
tail -2 /tmp/err.txt
done
