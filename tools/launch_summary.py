"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total ms and share."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            n = re.sub(r"\(.*", "", d["Kernel Name"])[:60]
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
            tot[n] += v; cnt[n] += 1
T = sum(tot.values())
print("%-60s %6s %10s %6s" % ("kernel", "count", "ms", "share"))
for n, v in tot.most_common(40):
    print("%-60s %6d %10.3f %5.1f%%" % (n, cnt[n], v, 100 * v / T))
print("%-60s %6d %10.3f" % ("total", sum(cnt.values()), T))
