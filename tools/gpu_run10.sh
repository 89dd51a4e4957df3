set -x
python bench.py > gpurun_out/bench_v33_full.log 2>&1; tail -1 gpurun_out/bench_v33_full.log | cut -c1-400
python bench.py --impl reference > gpurun_out/bench_ref_v33.log 2>&1; tail -1 gpurun_out/bench_ref_v33.log | cut -c1-300
