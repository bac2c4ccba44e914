# C4 solver time for forced CTA budgets of the L2-transport grid solver.
for n in 0 128 148 176 208 240 296; do
  FB_GRID_CTAS=$n python bench.py --config C4 --streams 1 --no-update --no-c4 --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4, FB_GRID_CTAS=$n: step %.1f us, solver %.1f us, frac %.3f' % (1e3*d['ms_per_step'], d['roofline']['launch_us'], d['roofline']['frac']))"
done
