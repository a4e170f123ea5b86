ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/loss_launches.csv python tools/loss_bench.py --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/loss_launches.csv')) if len(r)>5]
st=next(i for i,r in enumerate(rows) if "Kernel Name" in r); h=rows[st]; ki=h.index("Kernel Name"); mi=h.index("Metric Value")
seq=[(r[ki][:80], float(r[mi].replace(",",""))/1e3) for r in rows[st+1:]]
# find the last occurrence of LossWEpi and print the 16 launches around one fwd+bwd
idx=[i for i,(k,v) in enumerate(seq) if "ids_to_i32" in k]
i0=idx[-4] if len(idx)>=4 else 0
tot=0
for k,v in seq[i0-3:i0+16]:
    print(f"{v:8.1f} us  {k}")
PY
python -m pytest tests/test_gpu_losses.py -q -k graph 2>&1 | tail -2
