# Variant 3 (grid-resident solver): parity tests, then A/B against variant 2 / 1.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_golden.py -q -x 2>&1 | tail -5
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
for S in 1 8; do
  timeout 200 python bench.py --streams $S --variant 2 --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "S=$S v2"; tail -2 /tmp/err.txt
  for B in 148 296; do
    FB_GRID_CTAS=$B timeout 200 python bench.py --streams $S --variant 3 --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "S=$S v3 budget=$B"; tail -2 /tmp/err.txt
  done
done
FB_GRID_CTAS=64 timeout 200 python bench.py --streams 1 --variant 3 --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "S=1 v3 budget=64"; tail -2 /tmp/err.txt
FB_GRID_CTAS=32 timeout 200 python bench.py --streams 1 --variant 3 --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "S=1 v3 budget=32"; tail -2 /tmp/err.txt
timeout 200 python bench.py --streams 15 --variant 3 --steps 100 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "S=15 v3"; tail -2 /tmp/err.txt
timeout 300 python bench.py --config C4 --streams 1 --variant 1 --steps 40 --warmup 5 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "C4 v1"; tail -2 /tmp/err.txt
for B in 148 296; do
FB_GRID_CTAS=$B timeout 300 python bench.py --config C4 --streams 1 --variant 3 --steps 40 --warmup 5 --no-single --no-cpu-baseline 2>/tmp/err.txt | line "C4 v3 budget=$B"; tail -2 /tmp/err.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nltgv2_grid -s 3 -c 1 -o gpurun_out/prof_grid_r1_v1 python bench.py --steps 6 --warmup 3 --variant 3 --no-single --no-cpu-baseline > gpurun_out/ncu_grid.log 2>&1
tail -2 gpurun_out/ncu_grid.log
