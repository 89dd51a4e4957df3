# kernel-variant experiments: every build_exp/lib_*.so through the bench on an 8 z x 250 r sample of cfg5 (COFFE_B200_LIB override)
for lib in build_exp/lib_*.so; do
  name=$(basename $lib .so)
  COFFE_B200_LIB=$PWD/$lib python bench.py --nz 8 --nsep 250 --steps 3 --warmup 2 --no-cpu-baseline --no-covariance > gpurun_out/exp_$name.log 2>&1
  python - "$name" <<'PY'
import sys, json
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/exp_{name}.log").read().strip().splitlines()[-1])
    print(name, "points/s", round(d["value"]), "di_ms", round(d["roofline"]["launch_ms"], 3), "si_ms", round(d["roofline"]["step"]["si_kernel_ms"], 3))
except Exception as e:
    print(name, "FAILED", e)
PY
done
