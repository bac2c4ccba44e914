# e2e_update: per-camera loop in C (fb_update_run) vs in Python (FB_BENCH_PY_LOOP=1), same box.
for i in 1 2; do
  python bench.py --no-c4 --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); u=d['e2e_update']
print('C loop     ', round(u['value']), round(u['single_stream']['value']))"
  FB_BENCH_PY_LOOP=1 python bench.py --no-c4 --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); u=d['e2e_update']
print('Python loop', round(u['value']), round(u['single_stream']['value']))"
done
