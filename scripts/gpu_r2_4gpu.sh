# Two-GPU run of both bench arms exactly as the driver launches them (torchrun, one rank per GPU).
mkdir -p gpurun_out
nvidia-smi -L | head -4
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 ) > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
tail -c 3000 gpurun_out/r2_bench_4gpu.json; tail -5 gpurun_out/r2_bench_4gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 10 --warmup 3 ) > gpurun_out/r2_bench_4gpu_reference.json 2> gpurun_out/r2_bench_4gpu_reference.err
tail -c 800 gpurun_out/r2_bench_4gpu_reference.json; tail -4 gpurun_out/r2_bench_4gpu_reference.err
