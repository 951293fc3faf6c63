#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS mnemonics per kernel of csrc/libadsel_b200.so (cuobjdump -sass), written to
profiles/ so that the opcode evidence for tcgen05 / TMA / REDUX / cp.async / cluster use is in the tree."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "exploring-diversity-based-active-learning-for-3d-object-detection-in-autonomous-driving_b200", "csrc",
                   "libadsel_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "CREDUX", "REDUX", "REDG", "RED.",
         "ATOMG", "ATOMS", "LDGSTS", "LDGDEPBAR", "UCGABAR", "CCTL", "MAPA", "DSETP", "DADD", "DMUL", "DFMA", "BAR.SYNC", "LDS", "STS",
         "LDG", "STG", "VIMNMX", "ST.E", "LD.E")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("adsel::", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w) or (w.endswith(".") and op.startswith(w[:-1] + ".")):
                    kernels[cur][w.rstrip(".")] += 1
    lines = ["# SASS mnemonic counts per kernel of libadsel_b200.so (sm_100a), `python tools/sass_opcodes.py`",
             "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCATOMSWS = TMEM alloc, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,",
             "# SYNCS = mbarrier, CREDUX = warp REDUX, LDGSTS = cp.async, UCGABAR = barrier.cluster, MAPA = distributed shared memory address", ""]
    for name, c in kernels.items():
        body = ", ".join(f"{k} {v}" for k, v in sorted(c.items()) if k != "_total" and v)
        lines.append(f"{name}: {c['_total']} instructions; {body}")
    text = "\n".join(lines) + "\n"
    dest = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    open(dest, "w").write(text)
    print(text[:3000])


if __name__ == "__main__":
    main()
