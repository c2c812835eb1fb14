"""Top SASS lines by stall samples from `ncu -i rep --page source --csv` output (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = rows[1]
si, ci, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "(Not Issued)" not in x]
data = []
for k, r in enumerate(rows[2:]):
    try:
        v = int(r[ci]); ex = int(r[ei])
    except Exception:
        continue
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stall_cols), reverse=True)[:2]
    data.append((v, k, r[si].strip()[:70], ex, top))
tot = sum(d[0] for d in data)
print("total samples", tot, "sass lines", len(data), "dyn warp-instr", sum(d[3] for d in data))
for v, k, s, ex, top in sorted(data, reverse=True)[:n]:
    print(f"{v:6d} {100*v/tot:5.1f}%  line {k:5d} exec {ex:8d}  {s:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
