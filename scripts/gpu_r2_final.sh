# Round-2 evidence pass: full GPU test suite, smoke(), both bench arms as the driver runs them, launch lists.
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference ) > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; tail -3 gpurun_out/r2_bench_reference_final.err
( time python bench.py ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -3 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference_final.json').read().strip().splitlines()[-1])
print('value %.0f (%.1f us/step) e2e %.0f (%.1f us/step) solver_us %.1f frac %.3f launches %d' % (d['value'], 1e3*d['ms_per_step'], d['e2e']['value'], 1e3*d['e2e']['ms_per_step'], d['roofline']['launch_us'], d['roofline']['frac'], d['gpu_launches']))
print('reference arm value %.0f e2e %.0f ; e2e ratio %.1f' % (r['value'], r['e2e']['value'], d['e2e']['value']/r['e2e']['value']))
u=d['e2e_update']; ru=r.get('e2e_update') or {}
print('e2e_update %.0f fps (single %.0f) vs cpu %.0f -> %.1fx' % (u['value'], u['single_stream']['value'], ru.get('value', float('nan')), u['value']/ru.get('value', float('nan'))))
print('C4', json.dumps(d['configs']['C4'])[:400])
print('single', json.dumps(d.get('single_stream'))[:300])
PY
bash scripts/gpu_r2_launches.sh > gpurun_out/r2_launches_final.log 2>&1; tail -30 gpurun_out/r2_launches_final.log | head -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-update --no-c4 > gpurun_out/r2_bench_under_ncu.log 2>&1; python scripts/launch_table.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launch_table.md 2>&1; head -14 gpurun_out/r2_bench_launch_table.md
