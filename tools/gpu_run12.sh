set -x
python __graft_entry__.py --smoke 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v34_2gpu.log 2>&1; tail -1 gpurun_out/bench_v34_2gpu.log | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_2gpu.log 2>&1; tail -1 gpurun_out/bench_ref_2gpu.log | cut -c1-200
