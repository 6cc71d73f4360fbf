"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms, share."""
import collections
import csv
import sys


def main(path, per_launch=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        short = name.split("(")[0].replace("lsdm::<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        grid = row.get("Grid Size", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ms
        seq.append((short, ms, grid))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:5d} {v[1]:10.3f} {v[1] / tot:7.3f}")
    print(f"total {tot:.3f} ms over {len(seq)} launches")
    if per_launch:
        for i, (k, ms, grid) in enumerate(seq):
            print(f"{i:4d} {k[:60]:60s} {ms:9.4f} ms grid {grid}")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
