# compute-sanitizer (memcheck, racecheck) over the smoke invocation of the hot path, then one ncu --set full capture of the SI kernel
set -x
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:si_items_smem -c 1 -f -o gpurun_out/si_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline --no-covariance > gpurun_out/ncu_si.log 2>&1
ncu -i gpurun_out/si_full.ncu-rep --page raw --csv > gpurun_out/si_raw.csv; ncu -i gpurun_out/si_full.ncu-rep --page source --csv --print-source sass > gpurun_out/si_sass.csv
