"""Aggregate warp-state samples of `ncu --page source --csv` output over SASS line ranges.
usage: ncu_source_regions.py file.csv start:end[:name] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci = h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "(Not Issued)" not in x]
body = rows[2:]
def num(x):
    try: return int(x)
    except Exception: return 0
tot = sum(num(r[ci]) for r in body)
for spec in sys.argv[2:]:
    parts = spec.split(":")
    a, b = int(parts[0]), int(parts[1]); name = parts[2] if len(parts) > 2 else spec
    sub = body[a:b]
    n = sum(num(r[ci]) for r in sub)
    agg = sorted(((sum(num(r[i]) for r in sub), h[i]) for i in stall_cols), reverse=True)
    print(f"{name:10s} lines {a}-{b}: {n} samples ({100*n/tot:.1f}%): " + ", ".join(f"{k[6:]}={100*v/max(n,1):.0f}%" for v, k in agg[:8]))
