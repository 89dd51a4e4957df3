# 8-GPU run of both bench arms, launched the way the driver launches them
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; tail -1 gpurun_out/bench_8gpu.log | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_8gpu.log 2>&1; tail -1 gpurun_out/bench_ref_8gpu.log | cut -c1-400
