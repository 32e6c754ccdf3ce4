"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel, launches and device time.
python tools/launch_summary.py gpurun_out/launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
cnt, tot = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iU], 1.0)
    name = r[iN].split("(")[0]
    cnt[name] += 1; tot[name] += v
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k, t in tot.most_common():
    print(f"{k[:60]:60s} {cnt[k]:8d} {t:10.1f} {t/cnt[k]:8.2f} {100*t/total:5.1f}%")
