"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (times are cold-cache, serialised)."""
import collections
import csv
import re
import sys


def main(path, show_seq=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gstvd::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        seq.append((name, row["Grid Size"], v))
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:11.1f} us {v[0]:6d} x {v[1] / v[0]:9.2f} us {100 * v[1] / tot:5.1f}%  {k[:80]}")
    for s in seq[:show_seq]:
        print(f"   {s[2]:9.1f} us {s[1]:>16} {s[0][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
