timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2_pytest_2gpu_d.log 2>&1
tail -3 gpurun_out/r2_pytest_2gpu_d.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --config5 0 --nm-size 96 > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err
tail -c 1500 gpurun_out/r2_bench_2gpu_b.json
tail -3 gpurun_out/r2_bench_2gpu_b.err
