# Two-GPU run of both bench arms exactly as the driver launches them (torchrun, one rank per GPU).
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -c 3000 gpurun_out/r2_bench_8gpu.json; tail -5 gpurun_out/r2_bench_8gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2_bench_8gpu_reference.json 2> gpurun_out/r2_bench_8gpu_reference.err
tail -c 800 gpurun_out/r2_bench_8gpu_reference.json; tail -4 gpurun_out/r2_bench_8gpu_reference.err
