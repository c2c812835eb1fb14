"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count/mean/share."""
import csv, sys, collections
def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")); u = r[ui]
        v_us = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1.0)
        out.append((r[ki], v_us))
    return out
if __name__ == "__main__":
    for path in sys.argv[1:]:
        rows = load(path); agg = collections.OrderedDict()
        for k, v in rows:
            agg.setdefault(k[:70], []).append(v)
        tot = sum(sum(v) for k, v in agg.items() if "at::" not in k)
        print(f"# {path}")
        for k, v in agg.items():
            if "at::" in k: continue
            print(f"  {k:70s} n={len(v):3d} mean={sum(v)/len(v):8.1f} us  last={v[-1]:8.1f} us  share={100*sum(v)/tot:5.1f}%")
