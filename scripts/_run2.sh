timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_c.json 2> gpurun_out/r2_bench_2gpu_c.err
echo "rc=$?"
tail -c 2500 gpurun_out/r2_bench_2gpu_c.json
tail -3 gpurun_out/r2_bench_2gpu_c.err
nvidia-smi --query-gpu=memory.used --format=csv
