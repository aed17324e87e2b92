"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
shares, plus the Jacobi launches of the first fit.  Usage: launch_list.py in.csv out.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    v = float(r[ix["Metric Value"]].replace(",", "")); unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    seq.append((name, us))
agg = collections.OrderedDict(); tot = 0.0
for n, u in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += u; tot += u
out = ["kernel,launches,total_ms,share_pct"]
for k, (c, u) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append("%s,%d,%.3f,%.2f" % (k.replace(",", ";"), c, u / 1e3, 100 * u / tot))
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:16]))
gi = [i for i, (n, _) in enumerate(seq) if n.startswith("gram_dmma")]
if len(gi) > 1:
    fit = seq[gi[0]:gi[1]]
    print("one fit: %d launches, %.2f ms" % (len(fit), sum(u for _, u in fit) / 1e3))
    print("jacobi us:", [round(u) for n, u in fit if n.startswith("jacobi")])
