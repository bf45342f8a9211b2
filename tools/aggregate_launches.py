#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> CSV of shares."""
import collections
import csv
import re
import sys


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        full = r[kn].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        m = re.match(r"([\w:]+)", full)
        name = m.group(1) if m else full[:40]
        name = name.replace("at::native::", "torch:")
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    lines = ["kernel,launches,total_ns,share"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k},{v[0]},{v[1]:.0f},{v[1] / tot:.4f}")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(f"total_ms,{tot / 1e6:.3f}")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
