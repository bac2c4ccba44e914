# e2e leg diagnosis: where do the 40 us between `value` and `e2e` go?
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-update --no-c4 --no-cpu-baseline --no-single | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1: value %.1f us  e2e %.1f us  sync %.1f us  solver %.1f us' % (1e3*d['ms_per_step'], 1e3*d['e2e']['ms_per_step'], 1e3*d['e2e_sync']['ms_per_step'], d['roofline']['launch_us']))"; }
run default
FB_BENCH_PD_ITERS=1 run pd_iters_1
FB_BENCH_NO_XOUT=1 run no_xout
FB_PIPE_SINGLE_STAGE=1 run single_stage
