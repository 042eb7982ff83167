"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (markdown table)."""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("| launches | total ms | share | kernel |\n|---:|---:|---:|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {t:.3f} | {100*t/tot:.1f}% | `{k}` |")
print(f"\nTotal {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches.")
