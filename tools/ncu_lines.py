#!/usr/bin/env python
"""Per-SOURCE-LINE view of an ncu capture: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv;
python tools/ncu_lines.py f.csv [steps] [min_inst_per_step]  -> executed warp-instructions and stall samples per line."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    cur_file, hdr, col = None, None, None
    lines = []          # (file, lineno, text, inst, samples)
    total_s = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            col = {n: i for i, n in enumerate(hdr)}
            continue
        if r[0] == "Function Name" or hdr is None or len(r) != len(hdr):
            continue
        if r[0] != "":      # a source line (aggregated over its SASS)
            try:
                inst = int(r[col["Instructions Executed"]] or 0)
                smp = int(r[col["# Samples"]] or 0)
            except ValueError:
                continue
            lines.append((cur_file, int(r[0]), r[1].strip(), inst, smp))
            total_s += smp
    tot_i = sum(l[3] for l in lines)
    print(f"total warp-instructions/step {tot_i / steps:.0f}, samples {total_s}")
    for f, ln, text, inst, smp in lines:
        if inst / steps >= thresh or smp * 100.0 / max(total_s, 1) >= 0.5:
            print(f"{f}:{ln:5d} inst/step {inst / steps:7.1f}  samples {100.0 * smp / max(total_s, 1):5.1f}%  {text[:110]}")


if __name__ == "__main__":
    main()
