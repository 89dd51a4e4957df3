set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:di_items_smem -c 1 -f -o gpurun_out/items_v32_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_v32.log 2>&1; tail -1 gpurun_out/ncu_full_v32.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:si_items_smem -c 1 -f -o gpurun_out/si_v32_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_si_v32.log 2>&1; tail -1 gpurun_out/ncu_full_si_v32.log
(time python -m pytest tests/test_integrals.py -m gpu -x -q) 2>&1 | tail -3
