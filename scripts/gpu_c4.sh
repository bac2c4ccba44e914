set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_update.py -q -x 2>&1 | tail -2
timeout 400 python bench.py --config C4 --streams 1 --steps 60 --warmup 5 --no-single --no-cpu-baseline > gpurun_out/bench_r1_c4.json 2> gpurun_out/bench_r1_c4.err; tail -2 gpurun_out/bench_r1_c4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c4.json'))
print('C4 value %.0f fps e2e %.0f solver_us %.1f frac %.3f variant %s'%(d['value'],d['e2e']['value'],d['roofline']['launch_us'],d['roofline']['frac'],d['config']['solver_variant']))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_dual_edges|k_primal_vertices' -s 400 -c 4 -o gpurun_out/prof_streaming_c4 python bench.py --config C4 --streams 1 --steps 4 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
