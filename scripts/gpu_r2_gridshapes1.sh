# Single-stream solver time for forced cluster sizes / CTA shapes of the grid-resident solver.
for cfg in "0 0" "16 512" "16 384" "16 320" "12 512" "14 384" "8 512"; do
  set -- $cfg
  FB_GRID_CLUSTER=$1 FB_GRID_THREADS=$2 python bench.py --streams 1 --no-update --no-c4 --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1 stream, cluster $1 threads $2: step %.1f us, solver %.1f us' % (1e3*d['ms_per_step'], d['roofline']['launch_us']))"
done
