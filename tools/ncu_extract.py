"""Turn gpurun_out/{launches.csv, prof_fp16x3.ncu-rep, prof_fp16.ncu-rep} into the tracked evidence
under profiles/: launch list (copied), per-capture raw metric CSVs, traffic.json (DRAM bytes per launch,
read by bench.py) and a markdown summary printed to stdout."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    try:
        v = float(r[mi].replace(",", "")) / 1e6
    except ValueError:
        continue
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| share of GPU time | total ms | launches | kernel |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"| {100 * t / tot:.1f} % | {t:.3f} | {n} | `{k[:100]}` |")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{TAG}_launches_bench_fp16x3.csv"))

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic"]
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
traffic = {}
for prec in ("fp16x3", "fp16"):
    rep = os.path.join(G, f"prof_{prec}.ncu-rep")
    if not os.path.isfile(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    m = {n: (r[2][i], r[1][i]) for i, n in enumerate(r[0])}
    with open(os.path.join(P, f"{TAG}_ncu_sweep_{prec}_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        for k in sorted(m):
            w.writerow([k, m[k][0], m[k][1]])
    print(f"\n**{prec}** (`{m['Kernel Name'][0][:90]}`)\n\n| metric | value |\n|---|---|")
    for k in WANT:
        if k in m:
            print(f"| `{k}` | {m[k][0]} {m[k][1]} |")
    rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * MULT[m["dram__bytes_read.sum"][1]]
    wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * MULT[m["dram__bytes_write.sum"][1]]
    traffic[prec] = rd + wr
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("\ntraffic.json:", traffic)
