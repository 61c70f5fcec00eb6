"""Top stall-sample SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((r[ia].strip(), int(r[isamp]), int(r[iex])))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data)
print("total samples", tot, "warp-instructions", sum(d[2] for d in data), "sass lines", len(data))
top = sorted(enumerate(data), key=lambda x: -x[1][1])[:top_n]
for i, (s, n, e) in sorted(top):
    print(f"{i:5d} {n:6d} {100 * n / max(tot, 1):5.1f}% ex={e:9d}  {s[:100]}")
