mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_golden.py -q -x 2>&1 | tail -3
line() {
python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$1 variant=%s ctas=%s value %.0f (%.3f ms) e2e %.0f e2e_sync %.0f solver_us %.1f frac %.2f'%(d['config']['solver_variant'][:12],d['config']['ctas_per_stream'],d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e_sync']['value'],d['roofline']['launch_us'],d['roofline']['frac']))
except Exception as e:
    print('$1 FAILED', e)
"
}
B="--steps 100 --no-single --no-cpu-baseline"
for T in 512 384 320; do
for S in 1 8 12; do
FB_GRID_THREADS=$T timeout 200 python bench.py --streams $S --variant 3 $B 2>/tmp/err.txt | line "S=$S v3 threads=$T"; tail -2 /tmp/err.txt
done
done
timeout 200 python bench.py --streams 8 --variant 3 $B 2>/tmp/err.txt | line "S=8 v3 auto"; tail -2 /tmp/err.txt
timeout 300 python bench.py --config C4 --streams 1 --variant 3 --steps 40 --warmup 5 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "C4 v3"; tail -2 /tmp/err.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -o gpurun_out/prof_grid_r1_v9 python bench.py --steps 6 --warmup 3 --variant 3 --no-single --no-cpu-baseline > gpurun_out/ncu_grid.log 2>&1
tail -2 gpurun_out/ncu_grid.log
