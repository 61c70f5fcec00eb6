"""Join an `ncu --csv --log-file` launch list (one row per metric per kernel) with the plan's op names
(tests/prof_forward.py B ops.tsv) into one CSV row per launch.
usage: python profiles/launch_list.py ncu_log.csv ops.tsv out.csv"""
import csv
import sys

log, ops, out = sys.argv[1:4]
rows = [r for r in csv.reader(open(log, errors="replace")) if len(r) > 10]
h = rows[0]
iid, ik, ib, ig, im, iv, iu = (h.index(n) for n in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Value", "Metric Unit"))
launches = {}
for r in rows[1:]:
    if r[iid] == "ID":
        continue
    d = launches.setdefault(int(r[iid]), {"kernel": r[ik], "block": r[ib], "grid": r[ig]})
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu].strip().lower()
    if r[im].startswith("gpu__time_duration"):  # normalise to microseconds
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
    elif unit in ("kbyte", "mbyte", "gbyte"):
        v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    d[r[im]] = v
names = [ln.rstrip("\n").split("\t") for ln in open(ops)]
ids = sorted(launches)
assert len(ids) == len(names), (len(ids), len(names))
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "op", "tile_cfg", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB", "tensor_active_pct", "sm_throughput_pct", "l2_MB"])
    for i, (k, (name, tc)) in enumerate(zip(ids, names)):
        d = launches[k]
        t = d.get("gpu__time_duration.sum", 0.0)
        w.writerow([i, name, tc, d["kernel"].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:60], d["grid"], d["block"], round(t, 2), round(d.get("dram__bytes_read.sum", 0) / 1e6, 1),
                    round(d.get("dram__bytes_write.sum", 0) / 1e6, 1), round(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0), 1),
                    round(d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0), 1), round(d.get("lts__t_bytes.sum", 0) / 1e6, 1)])
print("wrote", out, len(ids), "launches")
