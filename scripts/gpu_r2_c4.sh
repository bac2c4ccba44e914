# C4 / L2-transport grid solver: parity tests, then the solver time.
timeout 600 python -m pytest tests/test_gpu_nltgv2.py -q -x 2>&1 | tail -3
python bench.py --config C4 --streams 1 --no-update --no-c4 --no-cpu-baseline --no-single 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4: step %.1f us, solver %.1f us, frac %.3f' % (1e3*d['ms_per_step'], d['roofline']['launch_us'], d['roofline']['frac']))"
python bench.py --no-update --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C2x8: step %.1f us, solver %.1f us; C4 block solver %.1f us frac %.3f' % (1e3*d['ms_per_step'], d['roofline']['launch_us'], d['configs']['C4']['roofline']['launch_us'], d['configs']['C4']['roofline']['frac']))"
