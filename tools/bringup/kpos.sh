python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -2
python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_small.csv python tools/gpu_diag.py time fp16x3 100000 1024 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_small.csv')) if len(r)>5]
st=next(i for i,r in enumerate(rows) if "Kernel Name" in r); h=rows[st]; ki=h.index("Kernel Name"); mi=h.index("Metric Value")
for r in rows[-6:]: print(r[ki][:70], float(r[mi].replace(",",""))/1e6)
PY
