set -x
python tools/quad_sweep.py '{"default": {}, "mu6": {"n_mu_panel_di": 6}}' 2>&1 | tail -2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:di_items_smem -c 1 -f -o gpurun_out/di_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_cfg5.log 2>&1; tail -1 gpurun_out/ncu_full_cfg5.log | cut -c1-300
python bench.py --no-cpu-baseline --no-covariance --no-extras 2>&1 | tail -1 | cut -c1-600
