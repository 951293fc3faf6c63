#!/usr/bin/env python
"""Offline reader of an ``ncu --set full --import-source on`` report: per kernel, the SASS instructions with the most
warp-stall samples and the samples by stall reason, as a small CSV for profiles/.

    python tools/ncu_hotspots.py gpurun_out/prof_greedy_r01c.ncu-rep profiles/r01c_greedy_hotspots.csv [--top 30] [--kernel substr]

Needs only the ``ncu`` CLI (no GPU): it shells out to ``ncu -i <rep> --page source --csv``.
"""
import argparse
import csv
import subprocess
import sys


def kernels_of(report):
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name, hdr, body, j = rows[i][1], rows[i + 1], [], i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            yield name, hdr, body
            i = j
        else:
            i += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--kernel", default="", help="only kernels whose name contains this")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    seen = set()
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        for name, hdr, body in kernels_of(args.report):
            if args.kernel not in name or not body or name in seen:      # first launch of every distinct kernel
                continue
            seen.add(name)
            si, src, addr = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Address")
            stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
            base, total = int(body[0][addr], 16), sum(int(r[si]) for r in body)
            w.writerow([f"# {name}: warp-stall samples per SASS instruction (ncu -i {args.report.split('/')[-1]} --page source "
                        f"--csv); top {args.top} of {len(body)} instructions, {total} samples. {args.note}"])
            w.writerow(["offset", "sass", "samples", "share_pct", "dominant_stall", "preceding_instruction"])
            for k in sorted(range(len(body)), key=lambda k: -int(body[k][si]))[: args.top]:
                r = body[k]
                dom = max(stall, key=lambda i: int(r[i] or 0))
                w.writerow([hex(int(r[addr], 16) - base), r[src].strip(), r[si], f"{100 * int(r[si]) / max(total, 1):.1f}",
                            f"{hdr[dom]}={r[dom]}", body[k - 1][src].strip() if k else ""])
            w.writerow([])
            w.writerow(["# all samples by stall reason"])
            agg = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall}
            for reason, v in sorted(agg.items(), key=lambda kv: -kv[1]):
                if v:
                    w.writerow([reason, v, f"{100 * v / max(sum(agg.values()), 1):.1f}"])
            w.writerow([])
    return 0


if __name__ == "__main__":
    sys.exit(main())
