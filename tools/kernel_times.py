"""Per-kernel average durations of one bench run from an ncu launch list (cold-cache, serialised: compare shares).

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file out.csv python bench.py --steps 2 --warmup 1
    python tools/kernel_times.py out.csv
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:8d} {t / 1000.0:12.1f} {t / 1000.0 / c:10.1f} {t / tot:7.1%}")


if __name__ == "__main__":
    main()
