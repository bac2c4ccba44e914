# Round 2: the driver's bench commands (both arms), outputs kept under gpurun_out/.
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt
( time python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -c 1500 gpurun_out/r2_bench_reference.json; tail -4 gpurun_out/r2_bench_reference.err
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f solver_us %.1f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['launch_us'], d['roofline']['frac']))
print(json.dumps(d.get('e2e_update'), indent=1)[:2500])
print('cpu_baseline', d.get('cpu_baseline'))
PY
tail -4 gpurun_out/r2_bench.err
