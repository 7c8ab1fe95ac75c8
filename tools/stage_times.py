"""Per-stage times of a decode step from a tools/timeline.py dump: stage = end-to-end delta between consecutive kernels."""
import collections
import sys

path = sys.argv[1]
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["qkv", "self", "o", "cq", "cross", "co", "f1", "f2"]
lines = open(path).read().split("\n")
i = [k for k, l in enumerate(lines) if l.startswith("one decode step")][0]
rows = [tuple(map(float, l.split()[:3])) for l in lines[i + 1:] if l.strip()]
ends = [s + d for s, d, _ in rows]
per = len(names)
layers = (len(rows) - 1) // per
acc = collections.defaultdict(list)
prev = ends[0]
for L in range(12):
    for j, n in enumerate(names):
        k = 1 + L * per + j
        acc[n].append(ends[k] - prev)
        prev = ends[k]
tot = 0.0
for n in names:
    m = sum(acc[n][1:]) / 11
    tot += m
    print(f"{n:6s} {m:6.2f} us")
print(f"layer  {tot:6.2f} us; tail stages:", [round(ends[k] - ends[k - 1], 1) for k in range(1 + 12 * per, len(rows))], "embed", round(ends[0] - rows[0][0], 1))
