#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE
step of bench.py (the launches between two consecutive k_build_text launches).
Usage: summarize_launches.py launches.csv [step_index]"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)", name)
    base = m.group(1) if m else name[:40]
    if base.endswith("_kernel") and "lambda" in name:
        # scan/select/for instantiations: say which host function they belong to
        owner = "esa_build" if "esa_build_device" in name else "anchor" if "anchor_queries_device" in name else "other"
        base += f"[{owner}]"
    return base


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "k_build_text" in n]
    a = starts[step]
    b = starts[step + 1] if step + 1 < len(starts) else len(rows)
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[a:b]:
        if "at::" in r["Kernel Name"]:
            continue  # torch's L2 flush between steps, not part of the step
        t = float(r["Metric Value"].replace(",", ""))
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        k = short(r["Kernel Name"])
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += t
        agg[k][1] += 1
        total += t
    print(f"# {path}: step {step}, {sum(c for _, c in agg.values())} launches, {total:.1f} us of kernel time")
    print(f"{'kernel':44s} {'launches':>8s} {'us':>10s} {'share':>7s}")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k[:44]:44s} {c:8d} {t:10.1f} {100 * t / total:6.1f}%")


if __name__ == "__main__":
    main()
