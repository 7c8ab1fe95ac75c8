"""Condenses a tools/ncu_summary.py text dump (one block per launch) into one line per kernel (means over its launches)."""
import collections
import re
import sys

KEYS = [("duration", "us"), ("grid", ""), ("regs/thread", ""), ("dyn smem/block", "KB"), ("dram read", "MB"), ("dram write", "MB"),
        ("crossbar -> SM read bytes", "MB x-bar->SM"), ("L2 hit rate %", "% L2 hit"), ("dram read % of ncu peak", "% dram"),
        ("tensor memory/MMA cycles active % (tcgen05)", "% tcgen05"), ("tensor pipe active % (mma.sync)", "% tensor pipe"),
        ("issue slots busy %", "% issue"), ("achieved occupancy %", "% occ"), ("fp64 pipe active %", "% fp64")]


def main(path):
    blocks, cur = [], None
    for l in open(path):
        if l.startswith("== "):
            name = re.sub(r"\(.*", "", l.split(": ", 1)[1]).replace("void ", "").replace("unnamed>::", "").replace("gstvd::", "").strip()
            cur = {"name": name}
            blocks.append(cur)
        elif cur is not None and l.strip():
            m = re.match(r"\s+(.*?)\s{2,}([-\d.eE+]+)\s*(\S*)", l)
            if m:
                v, unit = float(m.group(2)), m.group(3)
                if m.group(1) in ("dram read", "dram write", "crossbar -> SM read bytes"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
                if m.group(1) == "duration":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
                cur[m.group(1)] = v
    agg = collections.OrderedDict()
    for b in blocks:
        key = (b["name"], int(b.get("grid", 0)))
        agg.setdefault(key, []).append(b)
    tot = sum(b.get("duration", 0) for b in blocks)
    print(f"{len(blocks)} launches, sum of (serialised) durations {tot:.1f} us")
    names = ["regs", "smem KB", "dram rd MB", "dram wr MB", "xbar->SM MB", "L2 hit %", "dram %", "tcgen05 %", "tensor %", "issue %", "occ %", "fp64 %"]
    hdr = f"{'kernel':46s} {'grid':>5} {'n':>3} {'us':>7} {'share':>6}  " + "  ".join(f"{u:>11s}" for u in names)
    print(hdr)
    for (name, grid), bs in agg.items():
        def mean(k):
            xs = [b[k] for b in bs if k in b]
            return sum(xs) / len(xs) if xs else float("nan")
        d = mean("duration")
        print(f"{name[:46]:46s} {grid:5d} {len(bs):3d} {d:7.2f} {100 * d * len(bs) / tot:5.1f}%  " + "  ".join(f"{mean(k):11.2f}" for k, u in KEYS[2:]))
    dr = sum(b.get("dram read", 0) + b.get("dram write", 0) for b in blocks)
    print(f"DRAM traffic of all {len(blocks)} launches: {dr:.1f} MB (read + write)")


if __name__ == "__main__":
    main(sys.argv[1])
