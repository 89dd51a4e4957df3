ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-covariance > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_tmp.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[4:12]: print(r[k][:40], r[v])
PY
python tools/quad_sweep.py 2>&1 | tail -1
