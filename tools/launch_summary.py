"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per-kernel count, mean, share."""
import csv, re, sys, collections
def main(path, last_n=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"ce::\(anonymous namespace\)::|void |ce::", "", name)
        rows.append((name, ns))
    if last_n:
        rows = rows[-last_n:]
    agg = collections.OrderedDict()
    for n, ns in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("%-78s %5s %10s %10s %6s" % ("kernel", "n", "mean_us", "total_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-78s %5d %10.1f %10.1f %5.1f%%" % (n[:78], c, t / c / 1e3, t / 1e3, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot / 1e3, len(rows)))
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
