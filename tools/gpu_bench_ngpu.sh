# usage: bash tools/gpu_bench_ngpu.sh N -- both bench arms at N GPUs, launched the way the driver launches them
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_ref_${N}gpu.log | cut -c1-300
