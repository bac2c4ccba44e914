timeout 600 python -m pytest tests/test_gpu_hotpath_step.py tests/test_gpu_update.py -q -x 2>&1 | tail -2
for S in 1 8; do timeout 200 python bench.py --streams $S --steps 200 --no-single --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('S=$S value %.0f (%.1f us) e2e %.0f (%.1f us, link frac %.2f) e2e_sync %.0f solver_us %.1f'%(d['value'],1e3*d['ms_per_step'],d['e2e']['value'],1e3*d['e2e']['ms_per_step'],d['e2e']['h2d_link_frac'],d['e2e_sync']['value'],d['roofline']['launch_us']))"; done
