#!/usr/bin/env python
"""Per-instruction view of an ``ncu --set full --import-source on`` capture, read offline.

usage: ncu -i X.ncu-rep --page source --csv > src.csv;  python tools/ncu_hotspots.py src.csv [steps] [top]
Prints the executed warp-instructions per step by opcode, the stall samples by reason and the hottest SASS sites.
"""
import csv
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {name: i for i, name in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    by_op, by_reason, total_samples, total_inst = Counter(), Counter(), 0, 0
    sites = []
    for k, r in enumerate(body):
        sass = r[col["Source"]].strip()
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        op = op.rstrip(";")
        ex = int(r[col["Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        by_op[op.split(".")[0]] += ex
        total_inst += ex
        total_samples += sm
        for n in stall_cols:
            by_reason[n] += int(r[col[n]] or 0)
        sites.append((sm, ex, k, sass, {n: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0)}))
    print(f"warp-instructions executed: {total_inst} = {total_inst / steps:.0f} per step; samples {total_samples}")
    print("by opcode (per step):", ", ".join(f"{o} {c / steps:.0f}" for o, c in by_op.most_common(24)))
    print("stall samples:", ", ".join(f"{n[6:]} {100 * c / max(total_samples, 1):.1f}%" for n, c in by_reason.most_common(10)))
    print(f"top {top} sites by samples:")
    for sm, ex, k, sass, why in sorted(sites, reverse=True)[:top]:
        w = ", ".join(f"{n[6:]} {c}" for n, c in sorted(why.items(), key=lambda t: -t[1])[:3])
        print(f"  #{k:5d} {100 * sm / max(total_samples, 1):5.1f}%  exec/step {ex / steps:6.1f}  {sass[:70]:70s} {w}")


if __name__ == "__main__":
    main()
