python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu.log 2>&1; tail -1 gpurun_out/bench_8gpu.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/bench_4gpu.log 2>&1; tail -1 gpurun_out/bench_4gpu.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log | cut -c1-200
python bench.py --no-cpu-baseline --no-covariance --no-extras > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-200
