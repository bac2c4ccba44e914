# A/B of nltgv2_grid.cuh: working tree vs scratch_old_grid.cuh (hot-path step + C4 block).
run() { python bench.py --no-update --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1: C2x8 step %.1f us solver %.1f us | C4 solver %.1f us' % (1e3*d['ms_per_step'], d['roofline']['launch_us'], d['configs']['C4']['roofline']['launch_us']))"; }
run new; run new
cp flame_ros_b200/csrc/nltgv2_grid.cuh /tmp/new_grid.cuh
cp scratch_old_grid.cuh flame_ros_b200/csrc/nltgv2_grid.cuh
python -m flame_ros_b200.build > /dev/null 2>&1 && echo rebuilt-old
run old; run old
cp /tmp/new_grid.cuh flame_ros_b200/csrc/nltgv2_grid.cuh
python -m flame_ros_b200.build > /dev/null 2>&1 && echo rebuilt-new
run new
