"""Extract the judged metrics from an .ncu-rep into a small CSV (run where ncu is installed)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        for i, n in enumerate(h):
            if n in WANT:
                w.writerow([name, n, units[i], r[i]])
