"""Print the metrics the summaries quote from an `ncu --page raw --csv` export (one kernel launch per row)."""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[h], rows[h + 1]
    print("==", path)
    for r in rows[h + 2:]:
        for w in WANT:
            if w in hdr:
                k = hdr.index(w)
                print("  %-66s %s %s" % (w, r[k][:110], units[k]))
